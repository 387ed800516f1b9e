// harness_eq01.cpp -- test infrastructure (see bxref.h).  Pulls the reference's
// own translation unit in so that its solver class template is reachable;
// the file is read from /root/reference/lib/src at build time, never copied.
#include "itm-solver-equalities-01.cpp"

#include "harness_solver.hpp"

namespace bxref {
BXREF_DEFINE(eq01, solver_equalities_01coeff)

void
optimize_report(double* first_feasible_seconds, int64_t* restarts)
{
    if (first_feasible_seconds)
        *first_feasible_seconds = g_first_feasible;
    if (restarts)
        *restarts = g_restarts;
}
}
