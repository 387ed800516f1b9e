// Preprocessor over the LP handle: the variable fixing the
// reference runs in front of the solver when solver_parameters::preprocessor
// is "all" -- baryonyx::preprocess / affect / split,
// lib/src/preprocessor.cpp:139-851,857-906 -- restated over bxg_lp's flat rows:
// one counter per row (its number of unassigned elements), a variable -> rows
// index in CSR form per constraint type, a stack of pending assignments with a
// membership flag instead of the reference's linear search.  Host code, once
// per problem, no device needed; the result is again a bxg_lp, merged with the
// requested pre_constraint_order, whose rows go to bxg_problem_create.
//
// Kept as the reference has it, because the shrunk problem must be the same:
//   * the order of the passes (bounds, equal / greater / less rows, stack
//     drained last in first out) decides which of two contradicting
//     assignments wins: the stack refuses a variable it already holds and the
//     first value assigned to a variable stays (:96-111, :365);
//   * a variable no live row uses takes the value its objective factor likes,
//     and TRUE when the objective does not name it (:338-358);
//   * rows left with at most one unassigned element are dropped, the others
//     keep their id and lose the assigned elements, their right-hand side
//     reduced by the factors of those set to one (:814-847);
//   * a row that cannot hold (no value of its last variable satisfies it) is
//     the reference's bx_reach(): an error here, BXG_E_INVALID.
// Not restated: the quadratic objective (the LP reader refuses it, §7) and the
// "preprocessed.lp" file the reference writes into the working directory as a
// side effect (:709-712).
#ifndef BXG_PREPROCESS_HPP
#define BXG_PREPROCESS_HPP

namespace {

namespace pp {

// which of the three constraint vectors of the reference a row belongs to, in
// the order the preprocessor visits them: equal, greater, less
inline int
type_of(const lp::row& r)
{
    return r.op == 0 ? 0 : (r.op > 0 ? 1 : 2);
}

class preprocessor
{
    const bxg_lp& pb;
    const int n;
    std::vector<int> rows_of_type[3]; // indices into pb.rows, file order
    std::vector<int> count;           // per row of pb.rows: elements not yet assigned
    // variable -> rows holding it (one entry per element), per type
    std::vector<int> var_ptr[3], var_rows[3];
    std::vector<signed char> value; // -1: not assigned
    int assigned = 0;
    std::vector<std::pair<int, bool>> stack;
    std::vector<char> on_stack;
    std::string error;

    void push(int variable, bool v)
    {
        if (on_stack[variable])
            return;
        on_stack[variable] = 1;
        stack.emplace_back(variable, v);
    }

    // reduce + reduce_{equal,greater,less}_constraint (:157-260): the one
    // unassigned variable of a row and the value the row forces on it
    void reduce(int r)
    {
        const lp::row& row = pb.rows[r];
        int result = row.value, remaining = -1;
        for (size_t i = 0; i != row.elements.size(); ++i) {
            const int v = row.elements[i].second;
            if (value[v] < 0) {
                if (remaining != -1 && error.empty())
                    error = "preprocessor: two unassigned variables in a row counted down to one "
                            "(the reference's assertion remaining_index == -1)";
                remaining = static_cast<int>(i);
            } else
                result -= row.elements[i].first * value[v];
        }
        if (remaining < 0)
            return;
        const int factor = row.elements[remaining].first, variable = row.elements[remaining].second;
        const int t = type_of(row);
        const bool zero = t == 0 ? 0 == result : (t == 1 ? 0 >= result : 0 <= result);
        const bool one = t == 0 ? factor == result : (t == 1 ? factor >= result : factor <= result);
        if (zero && one)
            return;
        if (!zero && !one) {
            if (error.empty())
                error = "preprocessor: a constraint cannot be satisfied (the reference's bx_reach())";
            return;
        }
        push(variable, one);
    }

    // affects (:360-432)
    void drain()
    {
        while (!stack.empty()) {
            const auto top = stack.back();
            stack.pop_back();
            on_stack[top.first] = 0;
            if (value[top.first] < 0) {
                value[top.first] = top.second ? 1 : 0;
                ++assigned;
            }
            for (int t = 0; t != 3; ++t) {
                for (int e = var_ptr[t][top.first]; e != var_ptr[t][top.first + 1]; ++e) {
                    const int r = var_rows[t][e];
                    if (count[r] <= 0)
                        continue;
                    if (--count[r] == 1) {
                        reduce(r);
                        count[r] = 0;
                    }
                }
            }
        }
    }

    static bool all_factors_positive(const lp::row& row)
    {
        for (const auto& e : row.elements)
            if (e.first <= 0)
                return false;
        return true;
    }

    static int sum_factors(const lp::row& row)
    {
        int s = 0;
        for (const auto& e : row.elements)
            s += e.first;
        return s;
    }

    // try_affect_bounded_variable (:441-453)
    void bounded_variables()
    {
        for (int i = 0; i != n; ++i)
            if (pb.vars[i].min == pb.vars[i].max)
                push(i, pb.vars[i].max != 0);
    }

    // try_affect_variable (:455-559)
    void forced_rows()
    {
        for (int t = 0; t != 3; ++t) {
            for (int r : rows_of_type[t]) {
                const lp::row& row = pb.rows[r];
                if (count[r] == 1) {
                    reduce(r);
                    count[r] = 0;
                    continue;
                }
                if (!all_factors_positive(row))
                    continue;
                if (t == 0) {
                    if (sum_factors(row) == row.value || row.value == 0) {
                        count[r] = 0;
                        for (const auto& e : row.elements)
                            push(e.second, row.value != 0);
                    }
                } else if (t == 1) {
                    if (sum_factors(row) == row.value) {
                        count[r] = 0;
                        for (const auto& e : row.elements)
                            push(e.second, true);
                    } else if (row.value == 0)
                        count[r] = 0;
                } else {
                    if (row.value <= 0) {
                        count[r] = 0;
                        for (const auto& e : row.elements)
                            push(e.second, false);
                    } else if (sum_factors(row) == row.value)
                        count[r] = 0;
                }
            }
        }
    }

    // is_unused_variable + try_remove_unused_variable (:264-276, :337-358)
    void unused_variables()
    {
        std::vector<double> cost(n, 0.0);
        std::vector<char> named(n, 0);
        for (const auto& o : pb.objective)
            if (!named[o.second]) { // the reference takes the first element naming the variable
                named[o.second] = 1;
                cost[o.second] = o.first;
            }
        for (int i = 0; i != n; ++i) {
            if (value[i] >= 0)
                continue;
            bool used = false;
            for (int t = 0; t != 3 && !used; ++t)
                for (int e = var_ptr[t][i]; e != var_ptr[t][i + 1]; ++e)
                    if (count[var_rows[t][e]] > 0) {
                        used = true;
                        break;
                    }
            if (used)
                continue;
            bool v = true;
            if (named[i])
                v = pb.maximize ? cost[i] > 0 : cost[i] < 0;
            value[i] = v ? 1 : 0;
            ++assigned;
        }
    }

    void begin()
    {
        std::fill(value.begin(), value.end(), static_cast<signed char>(-1));
        std::fill(on_stack.begin(), on_stack.end(), 0);
        assigned = 0;
        stack.clear();
        error.clear();
        for (size_t r = 0; r != pb.rows.size(); ++r)
            count[r] = static_cast<int>(pb.rows[r].elements.size());
    }

    // make_problem (:651-717) with objective_function_exclude_copy,
    // variables_exclude_copy, constraints_exclude_copy
    std::unique_ptr<bxg_lp> shrink() const
    {
        std::unique_ptr<bxg_lp> out(new bxg_lp);
        out->maximize = pb.maximize;
        out->constant = pb.constant;
        std::vector<int> index(n, -1);
        out->vars.reserve(n - assigned);
        out->affected = pb.affected;
        out->affected.reserve(pb.affected.size() + assigned);
        for (int i = 0; i != n; ++i) {
            if (value[i] < 0) {
                index[i] = static_cast<int>(out->vars.size());
                out->vars.push_back(pb.vars[i]);
            } else
                out->affected.emplace_back(pb.vars[i].name, value[i] != 0);
        }
        for (const auto& o : pb.objective) {
            if (index[o.second] < 0) {
                if (value[o.second])
                    out->constant += o.first;
            } else
                out->objective.emplace_back(o.first, index[o.second]);
        }
        for (size_t r = 0; r != pb.rows.size(); ++r) {
            if (count[r] <= 1)
                continue;
            const lp::row& row = pb.rows[r];
            out->rows.emplace_back();
            lp::row& c = out->rows.back();
            c.id = row.id;
            c.op = row.op;
            c.value = row.value;
            for (const auto& e : row.elements) {
                if (index[e.second] >= 0)
                    c.elements.emplace_back(e.first, index[e.second]);
                else if (value[e.second])
                    c.value -= e.first;
            }
        }
        return out;
    }

public:
    explicit preprocessor(const bxg_lp& pb_)
      : pb(pb_)
      , n(static_cast<int>(pb_.vars.size()))
      , count(pb_.rows.size())
      , value(pb_.vars.size())
      , on_stack(pb_.vars.size())
    {
        for (size_t r = 0; r != pb.rows.size(); ++r)
            rows_of_type[type_of(pb.rows[r])].push_back(static_cast<int>(r));
        for (int t = 0; t != 3; ++t) {
            var_ptr[t].assign(n + 1, 0);
            for (int r : rows_of_type[t])
                for (const auto& e : pb.rows[r].elements)
                    ++var_ptr[t][e.second + 1];
            for (int i = 0; i != n; ++i)
                var_ptr[t][i + 1] += var_ptr[t][i];
            var_rows[t].resize(var_ptr[t][n]);
            std::vector<int> at(var_ptr[t].begin(), var_ptr[t].end() - 1);
            for (int r : rows_of_type[t])
                for (const auto& e : pb.rows[r].elements)
                    var_rows[t][at[e.second]++] = r;
        }
    }

    // operator()() (:621-640): preprocess
    std::unique_ptr<bxg_lp> run(std::string& err)
    {
        begin();
        bounded_variables();
        forced_rows();
        drain();
        unused_variables();
        drain();
        err = error;
        return err.empty() ? shrink() : nullptr;
    }

    // operator()(variable, value) (:601-619): affect, one half of split
    std::unique_ptr<bxg_lp> run(int variable, bool v, std::string& err)
    {
        begin();
        push(variable, v);
        drain();
        unused_variables();
        drain();
        err = error;
        return err.empty() ? shrink() : nullptr;
    }
};

inline int
finish(std::unique_ptr<bxg_lp> p, const std::string& err, int pre_order, bxg_lp** out)
{
    if (!p)
        return fail(BXG_E_INVALID, err);
    const std::string merr = lp_merge(*p, pre_order);
    if (!merr.empty())
        return fail(BXG_E_INVALID, "preprocessor: " + merr);
    *out = p.release();
    return BXG_OK;
}

} // namespace pp

} // namespace

extern "C" {

int
bxg_lp_preprocess(const bxg_lp* lp, int pre_order, bxg_lp** out)
{
    if (!lp || !out)
        return fail(BXG_E_INVALID, "bxg_lp_preprocess: bad argument");
    try {
        pp::preprocessor p(*lp);
        std::string err;
        auto shrunk = p.run(err);
        return pp::finish(std::move(shrunk), err, pre_order, out);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "bxg_lp_preprocess: out of host memory");
    } catch (...) {
        return fail(BXG_E_INVALID, "bxg_lp_preprocess: unexpected failure");
    }
}

int
bxg_lp_affect(const bxg_lp* lp, int variable, int value, int pre_order, bxg_lp** out)
{
    if (!lp || !out)
        return fail(BXG_E_INVALID, "bxg_lp_affect: bad argument");
    // the reference's precondition (preprocessor.cpp:866,884)
    if (variable < 0 || variable >= static_cast<int>(lp->vars.size()))
        return fail(BXG_E_INVALID, "bxg_lp_affect: variable index out of range");
    try {
        pp::preprocessor p(*lp);
        std::string err;
        auto shrunk = p.run(variable, value != 0, err);
        return pp::finish(std::move(shrunk), err, pre_order, out);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "bxg_lp_affect: out of host memory");
    } catch (...) {
        return fail(BXG_E_INVALID, "bxg_lp_affect: unexpected failure");
    }
}

int
bxg_lp_split(const bxg_lp* lp, int variable, int pre_order, bxg_lp** out_true, bxg_lp** out_false)
{
    if (!out_true || !out_false)
        return fail(BXG_E_INVALID, "bxg_lp_split: bad argument");
    *out_true = *out_false = nullptr;
    int rc = bxg_lp_affect(lp, variable, 1, pre_order, out_true);
    if (rc != BXG_OK)
        return rc;
    rc = bxg_lp_affect(lp, variable, 0, pre_order, out_false);
    if (rc != BXG_OK) {
        bxg_lp_destroy(*out_true);
        *out_true = nullptr;
    }
    return rc;
}

int
bxg_lp_affected_count(const bxg_lp* lp, int* count)
{
    if (!lp || !count)
        return fail(BXG_E_INVALID, "bxg_lp_affected_count: bad argument");
    *count = static_cast<int>(lp->affected.size());
    return BXG_OK;
}

const char*
bxg_lp_affected(const bxg_lp* lp, int i, int* value)
{
    if (!lp || i < 0 || i >= static_cast<int>(lp->affected.size()))
        return nullptr;
    if (value)
        *value = lp->affected[i].second ? 1 : 0;
    return lp->affected[i].first.c_str();
}

} // extern "C"

#endif
