// harness_ineq101.cpp -- test infrastructure (see bxref.h).  Pulls the reference's
// own translation unit in so that its solver class template is reachable;
// the file is read from /root/reference/lib/src at build time, never copied.
#include "itm-solver-inequalities-101.cpp"

#include "harness_solver.hpp"

namespace bxref {
BXREF_DEFINE(ineq101, solver_inequalities_101coeff)
}
