// bxg_kernels.cuh -- device data layout and kernel entry points.
#ifndef BXG_KERNELS_CUH
#define BXG_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace bxg {

enum
{
    FAM_01 = 0,
    FAM_101 = 1,
    FAM_Z = 2
};

enum
{
    ORDER_NONE = 0,
    ORDER_REVERSING,
    ORDER_RANDOM_SORTING,
    ORDER_INFEASIBILITY_DECR,
    ORDER_INFEASIBILITY_INCR,
    ORDER_LAGRANGIAN_DECR,
    ORDER_LAGRANGIAN_INCR,
    ORDER_PI_SIGN_CHANGE,
    ORDER_CYCLE
};

enum
{
    Z_BRANCH_AND_BOUND = 0,
    Z_EXHAUSTIVE,
    Z_UNCONSTRAINED_BRANCH_AND_BOUND,
    Z_UNCONSTRAINED_EXHAUSTIVE,
    Z_LINEAR
};

// What one constraint update needs to know about its row, in one 32-byte load.
struct __align__(16) RowHdr
{
    int r0;     // first element (row_ptr[k])
    int r_size; // number of elements
    int g0;     // first expanded column entry (row_gptr[k])
    int S;      // number of expanded column entries
    int bmin, bmax;
    int c_size; // elements with a negative factor (101)
    float bytes; // algorithmic bytes of one update
};

// Read-only structure shared by all lanes.  P lives in CSC order (see
// DESIGN.md): the gather for column j reads P[col_ptr[j] .. col_ptr[j+1])
// contiguously, and the row's own entry is one of those.
struct DevProblem
{
    int m, n, nnz, max_row;
    int p95_row; // 95th percentile of the row lengths
    int max_col; // longest column
    int max_S;   // most column entries met by one row (sum of its columns' lengths)
    int family, maximize;
    int xw32; // 32-bit words per x vector = 2 * (n/64 + 1)
    int n_obj;
    int obj_any_order; // objective sums are exact in any order
    int need_z;
    int tail_rows; // rows whose select_variables scan can read R[r_size]
    int max_zrow;  // (Z) longest row handled by branch-and-bound / exhaustive
    int max_amount; // largest amount by which a row can be violated
    double cost_constant;

    const RowHdr* row_hdr; // [m]
    const int* row_ptr;    // [m+1]
    const int* row_col;    // [nnz] column of each row element
    const int* row_cscpos; // [nnz] position of the element in CSC order
    const int4* row_info;  // [nnz] {col_ptr[j], col_ptr[j+1], CSC position, j} of each row element
    const int* row_a;      // [nnz] factor (101/Z) or nullptr
    const int* col_ptr;    // [n+1]
    const int* csc_row;    // [nnz] row of each column entry, ascending
    const int* csc_a;      // [nnz] factor in CSC order (101/Z) or nullptr
    const int* csc_rs;     // [nnz] row << 1 | (factor < 0), CSC order
    // column entries of every row, expanded: for row k the entries of its
    // columns one after the other, gp[row_gptr[k] + row_eptr[e] + t] =
    // {CSC position, row << 1 | (factor < 0)} of entry t of row element e
    const int2* gp;
    const int* row_gptr; // [m+1]
    const int* row_eptr; // [nnz] start of each row element inside its row's expansion
    const int* bmin;       // [m]
    const int* bmax;       // [m]
    const int* row_neg;    // [m] number of negative factors (101) or nullptr
    const float* row_bytes; // [m] algorithmic bytes of one update
    const int* ztype;      // [m] (Z) or nullptr
    const int* ex_start;   // [m] offset into ex_flat (Z)
    const int* ex_count;   // [m] number of patterns (Z)
    const int* ex_flat;    // feasible 0/1 patterns of the row as bit masks (Z)
    const int* obj_var;    // [n_obj]
    const double* obj_factor; // [n_obj]
};

// Per-batch parameters (uniform over lanes).
struct DevParams
{
    double theta, delta;
    double kappa_min, kappa_step, kappa_max;
    double pushing_k_factor, pushing_objective_amplifier;
    double w; // clamped: a sweep count
    long long limit;
    int pushes_limit, pushing_iteration_limit;
    int order; // constraint_order as configured (may be ORDER_CYCLE)
    double tail_value;
    // time_limit (itm-solver-common.hpp:228-232, stop_task of the optimizer,
    // itm-optimizer-common.hpp:857): value of %globaltimer after which a lane
    // stops at its next sweep boundary; 0 = no limit
    unsigned long long deadline_ns;
    const double* cost;   // [n] normalised costs
    const double* powtab; // [m+1] pow(k/m, alpha), computed by glibc on the host
};

// Per-lane control block.
struct LaneCtl
{
    uint32_t rng;
    int order_state; // current order (cycles when order == ORDER_CYCLE)
    int oR_len;
    int mo_len;
    int remaining; // result of the last sweep
    // results of solve
    int status;
    int best_remaining;
    int best_has_x;
    long long best_loop;
    double best_value;
    double objective;
    // optimizer rounds: solutions / advances offered to the population and how
    // many of them it took (try_update / try_advance)
    unsigned long long offers;
    unsigned long long inserts;
    unsigned long long updates;
    unsigned long long sweeps;
    double alg_bytes;
    unsigned long long n_path[4];   // updates by selection path (see Lane)
    unsigned long long cyc_path[4]; // cycles of thread 0 inside them
    unsigned long long cyc_phase[4]; // ... by phase: gather, sums, ties, affect
};

// The optimizer's population on the device: best_solution_recorder / storage
// (itm-optimizer-common.hpp:93-594).  Shared by the lanes of a batch (its
// workers); an insert holds `lock`, like the reference's writers hold
// m_indices_mutex.
struct DevPop
{
    int size;                 // init_population_size
    int victim_lo, victim_hi; // bad_solution_choose's range (:83-84)
    uint32_t* x;              // [size][xw32] members' vectors (m_data[i].x)
    double* value;            // [size]
    int* remaining;           // [size]
    unsigned long long* hash; // [size]
    long long* loop;          // [size]
    double* duration;         // [size] seconds on the device clock since the population was built
    int* order;               // [size] m_indices
    int* lock;
    const uint32_t* bastert;  // [xw32] m_bastert
    unsigned long long* first_ns; // device-clock time of the first solution a lane offered
    unsigned long long t0_ns;
};

// What best_solution_recorder::reinit (:528-554) decided for one lane: the
// scalar draws (normal distributions: glibc's log) are made on the host, the
// work over the n variables runs on the device (k_reinit).
struct ReinitPlan
{
    double kappa;     // kappa the restart begins with
    double var_p;     // mutation: probability to touch a variable
    double val_p;     // mutation: probability of the value 1
    uint32_t rng;     // engine state when the device takes over (start of the mutation loop)
    int mode;         // 0 keep x, 1 crossover(first, second), 2 member `first` := bastert, 3 := random
    int first;        // member id
    int second;       // member id, -1 = bastert, -2 = the lane's random vector
    int mutate;       // the per-variable mutation loop runs
};

// Per-batch lane storage (struct of arrays over lanes).
struct DevBatch
{
    int n_lanes;
    int order_cap; // entries per lane in oR / mo
    double* P;       // [lanes][nnz]  CSC order
    double* pi;      // [lanes][m]
    uint32_t* x;     // [lanes][xw32]
    uint32_t* bestx; // [lanes][xw32]
    int* act;        // [lanes][m] row activities of x
    double* Rv;      // [lanes][max_row + 2] values left in the slots of the reference's R
    int* oR;         // [lanes][order_cap]
    int2* mo;        // [lanes][order_cap]
    LaneCtl* ctl;    // [lanes]
    double* kappa;   // [lanes] input of bxg_batch_sweep
    // [lanes][5] theta, delta, kappa_min, kappa_step, init_policy_random of each
    // lane (parameter sweeps), or nullptr: the batch's values for every lane
    const double* lane_par;
    // row scratch when it does not fit the shared-memory budget, else nullptr
    unsigned char* g_scratch; // [lanes][lane_scratch_bytes]
    // (Z) items / node stack / subvector pool of the sub-solvers
    unsigned char* z_scratch; // [lanes][z_bytes]
    size_t z_bytes;
    // optimizer rounds (bxg_optimize_round)
    DevPop pop;
    const ReinitPlan* plan;            // [lanes]
    const unsigned long long* masks;   // [lanes][xw32 / 2] crossover_dist draws of a crossover
    const unsigned long long* aux;     // [lanes][xw32 / 2] init_with_random vector of a crossover
};

struct LaunchCfg
{
    int threads;
    size_t smem;
    int pi_in_smem, x_in_smem, r_in_smem, mo_in_smem, z_in_smem, exact_in_smem;
    int sort_cap;
    int stage_cap;
};

LaunchCfg plan_launch(const DevProblem& pb, int device, int n_lanes, int pairs_order);
size_t lane_scratch_bytes(const DevProblem& pb, const LaunchCfg& cfg);
size_t z_scratch_bytes(int max_zrow);

// The kernels are compiled once per solver family (bxg_kernels.cu with
// -DBXG_FAM=0|1|2); name_f0 / _f1 / _f2 are the entry points of the three
// translation units, the inline functions below pick by DevProblem::family.
#define BXG_DECLARE_FAMILY(F)                                                                    \
    void launch_begin_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,         \
                           const uint32_t* d_seeds, const uint32_t* d_x0, const uint32_t* d_rng, \
                           int reset_solver, int init_on_device, double init_random,             \
                           const double* d_cost_orig, const LaunchCfg& cfg, cudaStream_t s);     \
    void launch_order_init_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,    \
                                int lane_first, int lane_count, const LaunchCfg& cfg,            \
                                cudaStream_t s);                                                 \
    void launch_sweep_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,         \
                           int push, double obj_amp, const LaunchCfg& cfg, cudaStream_t s);      \
    void launch_solve_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,         \
                           const LaunchCfg& cfg, cudaStream_t s, int optimizer);                 \
    void launch_objective_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,     \
                               const LaunchCfg& cfg, cudaStream_t s);                            \
    void launch_optimize_f##F(const DevProblem& pb, const DevParams& pr, const DevBatch& b,      \
                              const LaunchCfg& cfg, cudaStream_t s);
BXG_DECLARE_FAMILY(0)
BXG_DECLARE_FAMILY(1)
BXG_DECLARE_FAMILY(2)
#undef BXG_DECLARE_FAMILY

#define BXG_BY_FAMILY(pb, name, ...)                                                             \
    do {                                                                                         \
        if ((pb).family == FAM_01)                                                               \
            name##_f0(__VA_ARGS__);                                                              \
        else if ((pb).family == FAM_101)                                                         \
            name##_f1(__VA_ARGS__);                                                              \
        else                                                                                     \
            name##_f2(__VA_ARGS__);                                                              \
    } while (0)

inline void
launch_begin(const DevProblem& pb, const DevParams& pr, const DevBatch& b, const uint32_t* d_seeds,
             const uint32_t* d_x0, const uint32_t* d_rng, int reset_solver, int init_on_device,
             double init_random, const double* d_cost_orig, const LaunchCfg& cfg, cudaStream_t s)
{
    BXG_BY_FAMILY(pb, launch_begin, pb, pr, b, d_seeds, d_x0, d_rng, reset_solver, init_on_device,
                  init_random, d_cost_orig, cfg, s);
}

inline void
launch_order_init(const DevProblem& pb, const DevParams& pr, const DevBatch& b, int lane_first,
                  int lane_count, const LaunchCfg& cfg, cudaStream_t s)
{
    BXG_BY_FAMILY(pb, launch_order_init, pb, pr, b, lane_first, lane_count, cfg, s);
}

inline void
launch_sweep(const DevProblem& pb, const DevParams& pr, const DevBatch& b, int push, double obj_amp,
             const LaunchCfg& cfg, cudaStream_t s)
{
    BXG_BY_FAMILY(pb, launch_sweep, pb, pr, b, push, obj_amp, cfg, s);
}

inline void
launch_solve(const DevProblem& pb, const DevParams& pr, const DevBatch& b, const LaunchCfg& cfg,
             cudaStream_t s, int optimizer = 0)
{
    BXG_BY_FAMILY(pb, launch_solve, pb, pr, b, cfg, s, optimizer);
}

inline void
launch_objective(const DevProblem& pb, const DevParams& pr, const DevBatch& b,
                 const LaunchCfg& cfg, cudaStream_t s)
{
    BXG_BY_FAMILY(pb, launch_objective, pb, pr, b, cfg, s);
}

// One call of optimize_functor::operator()'s loop body on every lane, after
// launch_reinit.
inline void
launch_optimize(const DevProblem& pb, const DevParams& pr, const DevBatch& b, const LaunchCfg& cfg,
                cudaStream_t s)
{
    BXG_BY_FAMILY(pb, launch_optimize, pb, pr, b, cfg, s);
}

#undef BXG_BY_FAMILY

} // namespace bxg

#endif
