// bxg_lp.hpp -- LP-file reader and constraint merging behind the C ABI
// (bxg_lp_*, include/bxg.h).  Included by bxg_api.cu; host code only.
//
// Produces, from LP text, exactly the arrays bxg_problem_create takes: what
// the reference obtains with make_problem (lib/src/parser.cpp:1065-1273),
// problem(raw_problem) (lib/src/problem.hpp:58-122) and
// make_merged_constraints (lib/src/merged-constraint.cpp:52-175,297-343).
// Same dialect, restated from its grammar: whitespace-separated chunks, a
// chunk starting with '\' comments out the rest of the line, chunks split at
// the separators < = > : - + [ ] * ^ and where a number ends; variables are
// numbered by first appearance (objective first); coefficients and right-hand
// sides are truncated to int; a variable repeated in a row has its factors
// added; rows over the same elements are merged into one [min, max] row.
// The file is tokenised once into memory (the reference re-reads an istream
// through a 10-token window, 512-byte chunks).
#ifndef BXG_LP_HPP
#define BXG_LP_HPP

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/stat.h>
#include <memory>
#include <fstream>
#include <string_view>
#include <unordered_map>

namespace {

namespace lp {

inline bool
same_word(std::string_view a, std::string_view b)
{
    if (a.size() != b.size())
        return false;
    for (size_t i = 0; i != a.size(); ++i)
        if (std::tolower(static_cast<unsigned char>(a[i])) != std::tolower(static_cast<unsigned char>(b[i])))
            return false;
    return true;
}

// Character classes of the dialect (parser.cpp: is_separator, is_char_name,
// is_number / number continuation), one table lookup per character.
enum : unsigned char { C_SPACE = 1, C_SEP = 2, C_NAME = 4, C_NUM = 8, C_NUM0 = 16 };

struct char_table
{
    unsigned char cls[256] = {};

    char_table()
    {
        for (int c = 0; c < 256; ++c) {
            unsigned char k = 0;
            if (std::isspace(c))
                k |= C_SPACE;
            if (std::strchr("<=>:-+[]*^", c) && c)
                k |= C_SEP;
            if ((c < 128 && std::isalnum(c)) || (c && std::strchr("!\"#$%&(),.;?@_{}~", c)))
                k |= C_NAME;
            if ((c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+')
                k |= C_NUM | C_NUM0;
            if (c == 'e' || c == 'E')
                k |= C_NUM;
            cls[c] = k;
        }
    }
};

inline const unsigned char*
classes()
{
    static const char_table t;
    return t.cls;
}

inline bool
is_separator(int c)
{
    return classes()[static_cast<unsigned char>(c)] & C_SEP;
}

inline bool
is_name_char(int c)
{
    return classes()[static_cast<unsigned char>(c)] & C_NAME;
}

inline bool
is_number_char(int c)
{
    return classes()[static_cast<unsigned char>(c)] & C_NUM;
}

inline bool
starts_number(int c)
{
    return classes()[static_cast<unsigned char>(c)] & C_NUM0;
}

inline bool
is_keyword(std::string_view s)
{
    static const char* const words[] = { "binary", "binaries", "bin", "bound", "bounds", "general", "generals",
                                         "gen", "end", "st", "subject", "sush", "s.t.", "st." };
    for (const char* w : words)
        if (same_word(w, s))
            return true;
    return false;
}

// The token stream of a whole file (views into `text`).
struct tokens
{
    std::string text;
    std::vector<std::string_view> tok;
    size_t pos = 0;

    explicit tokens(std::string t)
      : text(std::move(t))
    {
        const char* s = text.data();
        const size_t n = text.size();
        const unsigned char* cls = classes();
        tok.reserve(n / 3 + 16);
        size_t i = 0;
        while (i < n) {
            while (i < n && (cls[static_cast<unsigned char>(s[i])] & C_SPACE))
                ++i;
            if (i == n)
                break;
            size_t e = i;
            while (e < n && !(cls[static_cast<unsigned char>(s[e])] & C_SPACE))
                ++e;
            if (s[i] == '\\') { // comment: the rest of the line
                while (e < n && s[e] != '\n')
                    ++e;
                i = e;
                continue;
            }
            // operator>> into char[512] cuts chunks at 511 characters
            for (size_t c0 = i; c0 < e; c0 += 511)
                split(c0, std::min(e, c0 + 511));
            i = e;
        }
    }

    void split(size_t b, size_t e)
    {
        const char* s = text.data();
        const unsigned char* cls = classes();
        size_t cur = b;
        while (cur < e) {
            const unsigned char k0 = cls[static_cast<unsigned char>(s[cur])];
            const size_t start = cur++;
            if ((k0 & C_SEP) || cur == e) {
                tok.emplace_back(s + start, 1);
                continue;
            }
            // a number stops where a number cannot go on, anything stops at a separator
            const unsigned char stop = C_SEP, need = (k0 & C_NUM0) ? C_NUM : 0;
            while (cur < e) {
                const unsigned char k = cls[static_cast<unsigned char>(s[cur])];
                if ((k & stop) || (k & need) != need)
                    break;
                ++cur;
            }
            tok.emplace_back(s + start, cur - start);
        }
    }

    std::string_view at(size_t k) const { return pos + k < tok.size() ? tok[pos + k] : std::string_view(); }
    void pop(size_t k = 1) { pos += k; }
};

inline bool
read_real(std::string_view s, double& v)
{
    if (s.size() >= 3 && same_word(s, "inf")) {
        v = std::numeric_limits<double>::infinity();
        return true;
    }
    if (s.empty() || s.size() > 320)
        return false;
    {
        // strtod can only succeed on [+-] digit . inf nan; plain digit strings
        // below 2^53 convert exactly without it
        const char c = s[0];
        if (!((c >= '0' && c <= '9') || c == '.' || c == '+' || c == '-' || c == 'i' || c == 'I' || c == 'n' || c == 'N'))
            return false;
        if (s.size() <= 15) {
            uint64_t acc = 0;
            size_t i = 0;
            for (; i != s.size() && s[i] >= '0' && s[i] <= '9'; ++i)
                acc = acc * 10 + static_cast<uint64_t>(s[i] - '0');
            if (i == s.size()) {
                v = static_cast<double>(acc);
                return true;
            }
        }
    }
    const std::string z(s);
    char* end = nullptr;
    const double r = std::strtod(z.c_str(), &end); // sscanf("%lf") of the reference
    if (end == z.c_str())
        return false;
    v = r;
    return true;
}

// a (possibly signed) coefficient over up to two tokens; `used` = tokens read
inline void
read_signed(std::string_view a, std::string_view b, double& v, int& used)
{
    double r;
    if (a == "-" || a == "+") {
        const double sign = a == "-" ? -1.0 : 1.0;
        if (read_real(b, r)) {
            v = sign * r;
            used = 2;
        } else {
            v = sign;
            used = 1;
        }
        return;
    }
    if (read_real(a, r)) {
        v = r;
        used = 1;
    } else {
        v = 1.0;
        used = 0;
    }
}

struct element
{
    double factor = 0;
    std::string_view name;
    int used = 0;
    bool ok = true;
};

inline bool
is_name(std::string_view s)
{
    for (char c : s)
        if (!is_name_char(c))
            return false;
    return true;
}

inline element
read_element(const tokens& t)
{
    element e;
    read_signed(t.at(0), t.at(1), e.factor, e.used);
    const std::string_view w = t.at(e.used);
    if (!is_keyword(w) && !w.empty() && is_name_char(w[0])) {
        if (!is_name(w)) {
            e.ok = false;
            return e;
        }
        e.name = w;
        e.used += 1;
    }
    return e;
}

// operator over up to two tokens: 0 none, else tokens read; op: -1 <=, 0 =, +1 >=
inline int
read_operator(std::string_view a, std::string_view b, int& op)
{
    if (a == "<") {
        op = -1;
        return b == "=" ? 2 : 1;
    }
    if (a == ">") {
        op = 1;
        return b == "=" ? 2 : 1;
    }
    if (a == "=") {
        if (b == "<") {
            op = -1;
            return 2;
        }
        if (b == ">") {
            op = 1;
            return 2;
        }
        op = 0;
        return b == "=" ? 2 : 1;
    }
    return 0;
}

// Allocator of the big look-up table: 2 MB alignment + MADV_HUGEPAGE, so that
// random probes of a table of tens of MB do not pay a TLB miss each.
template<typename T>
struct huge_allocator
{
    using value_type = T;
    huge_allocator() = default;
    template<typename U>
    huge_allocator(const huge_allocator<U>&) {}
    T* allocate(size_t n)
    {
        constexpr size_t huge = size_t(2) << 20;
        size_t bytes = n * sizeof(T);
        void* p = nullptr;
        if (bytes >= huge) {
            bytes = (bytes + huge - 1) / huge * huge;
            if (posix_memalign(&p, huge, bytes) != 0)
                throw std::bad_alloc();
            madvise(p, bytes, MADV_HUGEPAGE);
        } else if (!(p = std::malloc(bytes ? bytes : 1)))
            throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { std::free(p); }
    template<typename U>
    bool operator==(const huge_allocator<U>&) const { return true; }
    template<typename U>
    bool operator!=(const huge_allocator<U>&) const { return false; }
};

// Variable names -> ids.  Open addressing; an entry carries the name itself
// (up to 19 characters inline, longer ones through `keys`) and the variable's
// (owner, position) stamp, so that a look-up touches one cache line; the
// parser prefetches the entries of the names a few elements ahead.
struct name_table
{
    struct entry
    {
        int32_t id = -1; // -1: free
        int32_t owner = -1, position = 0;
        uint8_t len = 0; // 255: longer than the inline room
        char name[19];
    };
    static_assert(sizeof(entry) == 32, "two entries per cache line");

    std::vector<entry, huge_allocator<entry>> slot;
    std::vector<std::string_view> keys; // id -> name (views into the token text)
    size_t mask = 0;

    static uint64_t hash(std::string_view s)
    {
        uint64_t h = 0x9e3779b97f4a7c15ull ^ s.size();
        const char* p = s.data();
        size_t n = s.size();
        while (n >= 8) {
            uint64_t w;
            std::memcpy(&w, p, 8);
            h = (h ^ w) * 0xff51afd7ed558ccdull;
            h ^= h >> 32;
            p += 8, n -= 8;
        }
        uint64_t w = 0;
        std::memcpy(&w, p, n);
        h ^= w;
        // full avalanche: names that differ in their last characters only
        // must not share the low (index) bits
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 33;
        h *= 0xc4ceb9fe1a85ec53ull;
        return h ^ (h >> 33);
    }

    void grow()
    {
        const size_t cap = slot.empty() ? 2048 : slot.size() * 2;
        std::vector<entry, huge_allocator<entry>> next(cap);
        mask = cap - 1;
        for (const entry& e : slot) {
            if (e.id < 0)
                continue;
            size_t i = hash(keys[e.id]) & mask;
            while (next[i].id >= 0)
                i = (i + 1) & mask;
            next[i] = e;
        }
        slot.swap(next);
    }

    void prefetch(std::string_view name) const
    {
        if (mask)
            __builtin_prefetch(&slot[hash(name) & mask]);
    }

    // entry of `name`; nullptr when absent and !insert; fresh set when inserted
    entry* find(std::string_view name, bool insert, bool& fresh)
    {
        fresh = false;
        if ((keys.size() + 1) * 2 > slot.size())
            grow();
        const size_t n = name.size();
        size_t i = hash(name) & mask;
        while (slot[i].id >= 0) {
            entry& e = slot[i];
            if (n <= sizeof(e.name) ? (e.len == n && std::memcmp(e.name, name.data(), n) == 0)
                                    : (e.len == 255 && keys[e.id] == name))
                return &e;
            i = (i + 1) & mask;
        }
        if (!insert)
            return nullptr;
        entry& e = slot[i];
        e.id = static_cast<int32_t>(keys.size());
        e.owner = -1;
        e.position = 0;
        if (n <= sizeof(e.name)) {
            e.len = static_cast<uint8_t>(n);
            std::memcpy(e.name, name.data(), n);
        } else
            e.len = 255;
        keys.push_back(name);
        fresh = true;
        return &e;
    }
};

struct row
{
    std::vector<std::pair<int, int>> elements; // (factor, variable)
    int value = 0;
    int id = 0;
    int op = 0;
};

struct variable
{
    std::string name;
    int min = 0, max = std::numeric_limits<int>::max();
    int type = 0; // variable_type: real, binary, general
};

} // namespace lp

} // namespace

struct bxg_lp
{
    int maximize = 0;
    double constant = 0;
    std::vector<std::pair<double, int>> objective; // (factor, variable), first-appearance order
    std::vector<lp::variable> vars;
    std::vector<lp::row> rows; // as parsed, ids in file order
    int problem_type = 0;
    // variables the preprocessor fixed (name, value): affected_variables, core:654-667
    std::vector<std::pair<std::string, bool>> affected;
    // merged constraints (CSR)
    std::vector<int32_t> row_ptr{ 0 }, col, factor, cmin, cmax;
};

namespace {

// parse(), parser.cpp:1065-1260.  Returns an empty string or the reference's
// file_format_error_tag name.
std::string
lp_parse(std::string text, bxg_lp& out)
{
    lp::tokens t(std::move(text));
    lp::name_table ids;
    // an entry's (owner, position): the variable's element in the objective
    // (owner -2) or in the row being read (owner = its id); else "not yet"
    auto variable = [&](std::string_view name) { // keys are views into t.text, alive for the whole parse
        bool fresh;
        lp::name_table::entry* e = ids.find(name, true, fresh);
        if (fresh)
            out.vars.push_back({ std::string(name) });
        return e;
    };
    auto known = [&](std::string_view name) {
        bool fresh;
        const lp::name_table::entry* e = ids.find(name, false, fresh);
        return e ? e->id : -1;
    };
    // names some elements ahead: bring their table entries into cache
    auto look_ahead = [&]() {
        const std::string_view a = t.at(24), b = t.at(25);
        if (a.size() > 1 || (!a.empty() && !lp::is_separator(a[0])))
            ids.prefetch(a);
        if (b.size() > 1 || (!b.empty() && !lp::is_separator(b[0])))
            ids.prefetch(b);
    };

    const std::string_view head = t.at(0);
    if (lp::same_word(head, "maximize") || lp::same_word(head, "maximum") || lp::same_word(head, "max"))
        out.maximize = 1;
    else if (lp::same_word(head, "minimize") || lp::same_word(head, "minimum") || lp::same_word(head, "min"))
        out.maximize = 0;
    else
        return "bad_objective_function_type";
    t.pop();
    if (!lp::is_keyword(t.at(0)) && lp::same_word(t.at(1), ":"))
        t.pop(2); // objective label

    while (!lp::is_keyword(t.at(0))) {
        if (t.at(0).empty())
            return "bad_objective";
        if (t.at(0) == "[" || ((t.at(0) == "+" || t.at(0) == "-") && t.at(1) == "["))
            return "quadratic objective (outside this engine)";
        const lp::element e = lp::read_element(t);
        if (!e.ok || e.used == 0) // nothing consumed: the reference reports bad_objective (or spins)
            return "bad_objective";
        if (e.name.empty())
            out.constant += e.factor;
        else {
            // (the reference looks a repeated variable up with a linear find_if,
            // parser.cpp:1085-1097: quadratic in the objective's length; a
            // per-variable position gives the same result in O(1))
            look_ahead();
            lp::name_table::entry* v = variable(e.name);
            if (v->owner != -2) {
                v->owner = -2;
                v->position = static_cast<int32_t>(out.objective.size());
                out.objective.emplace_back(e.factor, v->id);
            } else
                out.objective[v->position].first += e.factor;
        }
        t.pop(e.used);
    }

    // subject to
    {
        int used = 0;
        const std::string_view a = t.at(0), b = t.at(1), c = t.at(2);
        if (lp::same_word(a, "st") || lp::same_word(a, "st.") || lp::same_word(a, "s.t") || lp::same_word(a, "s.t."))
            used = lp::same_word(b, ":") ? 2 : 1;
        else if ((lp::same_word(a, "subject") && lp::same_word(b, "to")) ||
                 (lp::same_word(a, "sush") && lp::same_word(b, "that")))
            used = lp::same_word(c, ":") ? 3 : 2;
        if (used) {
            t.pop(used);
            int next_id = 0;
            while (!t.at(0).empty() && !lp::is_keyword(t.at(0))) {
                lp::row r;
                if (lp::is_name_char(t.at(0)[0]) && t.at(1) == ":")
                    t.pop(2); // constraint label
                bool first = true;
                while (first || (!t.at(0).empty() && !(t.at(0)[0] == '<' || t.at(0)[0] == '>' || t.at(0)[0] == '='))) {
                    first = false;
                    const lp::element e = lp::read_element(t);
                    if (!e.ok || e.used == 0)
                        return "bad_constraint";
                    look_ahead();
                    lp::name_table::entry* v = variable(e.name);
                    if (v->owner != next_id) {
                        v->owner = next_id;
                        v->position = static_cast<int32_t>(r.elements.size());
                        r.elements.emplace_back(static_cast<int>(e.factor), v->id);
                    } else
                        r.elements[v->position].first += static_cast<int>(e.factor);
                    t.pop(e.used);
                }
                const int opn = lp::read_operator(t.at(0), t.at(1), r.op);
                if (!opn)
                    return "bad_constraint";
                t.pop(opn);
                double v;
                int vn;
                lp::read_signed(t.at(0), t.at(1), v, vn);
                if (vn == 0) // no right-hand side: the reference's pop_front(0) contract aborts (parser.cpp:302)
                    return "bad_constraint";
                r.value = static_cast<int>(v);
                t.pop(vn);
                r.id = next_id++;
                out.rows.push_back(std::move(r));
            }
        }
    }

    // bounds
    if (lp::same_word(t.at(0), "bounds") || lp::same_word(t.at(0), "bound")) {
        t.pop(lp::same_word(t.at(1), ":") ? 2 : 1);
        while (!lp::is_keyword(t.at(0))) {
            if (t.at(0).empty())
                return "bad_bound";
            double lo = 0.0, hi = std::numeric_limits<double>::infinity();
            std::string_view name;
            int used = 0;
            auto right = [&](size_t k, double& value, int& n) { // operator [sign] number
                int op;
                const int on = lp::read_operator(t.at(k), t.at(k + 1), op);
                if (!on)
                    return false;
                size_t p = k + on;
                double sign = 1.0;
                if (t.at(p) == "+" || t.at(p) == "-") {
                    sign = t.at(p) == "-" ? -1.0 : 1.0;
                    ++p;
                }
                double r;
                if (!lp::read_real(t.at(p), r))
                    return false;
                value = sign * r;
                n = static_cast<int>(p + 1 - k);
                return true;
            };
            const std::string_view a = t.at(0);
            const bool numeric = !a.empty() && (lp::is_number_char(a[0]) || lp::same_word(a, "infinity") || lp::same_word(a, "inf"));
            if (numeric) {
                size_t p = 0;
                double sign = 1.0;
                if (a == "+" || a == "-") {
                    sign = a == "-" ? -1.0 : 1.0;
                    p = 1;
                }
                double r;
                if (!lp::read_real(t.at(p), r))
                    return "bad_bound";
                lo = sign * r;
                int op;
                const int on = lp::read_operator(t.at(p + 1), t.at(p + 2), op);
                if (!on)
                    return "bad_bound";
                p += 1 + on;
                if (!lp::is_name(t.at(p)) || t.at(p).empty())
                    return "bad_bound";
                name = t.at(p);
                used = static_cast<int>(p + 1);
                if (!t.at(used).empty() && (t.at(used)[0] == '<' || t.at(used)[0] == '>' || t.at(used)[0] == '=')) {
                    int n = 0;
                    if (!right(used, hi, n) || lo > hi)
                        return "bad_bound";
                    used += n;
                } else
                    used += 1; // the reference consumes one token too many here (parser.cpp:924-929)
            } else if (!a.empty() && lp::is_name_char(a[0])) {
                if (!lp::is_name(a))
                    return "bad_bound";
                name = a;
                used = 1;
                if (!t.at(1).empty() && (t.at(1)[0] == '<' || t.at(1)[0] == '>' || t.at(1)[0] == '=')) {
                    int n = 0;
                    if (!right(1, hi, n))
                        return "bad_bound";
                    lo = 0.0;
                    used += n;
                } else {
                    lo = -std::numeric_limits<double>::infinity();
                }
            } else
                return "bad_bound";
            const int id = known(name);
            if (id < 0)
                return "bad_bound";
            out.vars[id].min = lo == -std::numeric_limits<double>::infinity() ? std::numeric_limits<int>::min()
                                                                              : static_cast<int>(lo);
            out.vars[id].max = hi == std::numeric_limits<double>::infinity() ? std::numeric_limits<int>::max()
                                                                             : static_cast<int>(hi);
            t.pop(used);
        }
    }

    if (lp::same_word(t.at(0), "binary") || lp::same_word(t.at(0), "binaries") || lp::same_word(t.at(0), "bin")) {
        t.pop(lp::same_word(t.at(1), ":") ? 2 : 1);
        while (!lp::is_keyword(t.at(0))) {
            // (truncated text: the reference's loop would not end here either; a
            // library entry point reports it instead)
            if (t.at(0).empty())
                return "bad_binary";
            const int id = known(t.at(0));
            if (id < 0)
                return "bad_binary";
            out.vars[id].type = 1;
            out.vars[id].min = 0;
            out.vars[id].max = 1;
            t.pop();
        }
    }
    if (lp::same_word(t.at(0), "general") || lp::same_word(t.at(0), "generals") || lp::same_word(t.at(0), "gen")) {
        t.pop(lp::same_word(t.at(1), ":") ? 2 : 1);
        while (!lp::is_keyword(t.at(0))) {
            if (t.at(0).empty())
                return "bad_general";
            const int id = known(t.at(0));
            if (id < 0)
                return "bad_general";
            out.vars[id].type = 2;
            t.pop();
        }
    }
    if (lp::same_word(t.at(0), "end")) {
        t.pop(lp::same_word(t.at(1), ":") ? 2 : 1);
        return t.at(0).empty() ? "" : "bad_end_of_file";
    }
    return "bad_end";
}

// problem::which_problem_type (problem.hpp:72-122) and
// make_merged_constraints for pre_order none / memory
// (merged-constraint.cpp:52-175,286-315).
std::string
lp_merge(bxg_lp& p, int pre_order)
{
    if (pre_order < -1 || pre_order > 11)
        return "unknown pre_constraint_order";
    int coef = 0;
    bool inequality = false;
    for (const lp::row& r : p.rows) {
        inequality = inequality || r.op != 0;
        for (const auto& e : r.elements) {
            if (e.first <= -2 || e.first >= 2)
                coef = 2;
            else if (e.first == -1 && coef == 0)
                coef = 1;
        }
    }
    p.problem_type = (inequality ? 3 : 0) + coef;
    if (pre_order == -1) { // BXG_LP_RAW: the parse as it is, for bxg_lp_preprocess (the reference merges after its preprocessor)
        p.row_ptr.assign(1, 0);
        p.col.clear(), p.factor.clear(), p.cmin.clear(), p.cmax.clear();
        return "";
    }

    struct merged
    {
        const std::vector<std::pair<int, int>>* elements;
        int min, max, id;
    };
    struct key_hash
    {
        size_t operator()(const std::vector<std::pair<int, int>>* v) const
        {
            size_t seed = v->size();
            for (const auto& e : *v)
                seed ^= static_cast<size_t>(e.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
            return seed;
        }
    };
    struct key_eq
    {
        bool operator()(const std::vector<std::pair<int, int>>* a, const std::vector<std::pair<int, int>>* b) const
        {
            return *a == *b;
        }
    };
    std::unordered_map<const std::vector<std::pair<int, int>>*, size_t, key_hash, key_eq> seen;
    std::vector<merged> out;
    // The order of the three passes decides which row of a merged group keeps
    // its id: equal, less, greater for none / memory / p1-p4
    // (make_unsorted_merged_constraints), the order named by the parameter for
    // the six type orders (make_ordered_merged_constraints, :177-210).
    static const int pass_order[12][3] = { { 0, -1, 1 }, { 0, -1, 1 }, { -1, 1, 0 }, { -1, 0, 1 },
                                           { 1, -1, 0 }, { 1, 0, -1 }, { 0, -1, 1 }, { 0, 1, -1 },
                                           { 0, -1, 1 }, { 0, -1, 1 }, { 0, -1, 1 }, { 0, -1, 1 } };
    const bool by_type = pre_order >= 2 && pre_order <= 7, special = pre_order >= 8;
    for (int pass : pass_order[pre_order]) {
        for (const lp::row& r : p.rows) {
            if (r.op != pass)
                continue;
            auto it = seen.find(&r.elements);
            if (it == seen.end()) {
                seen.emplace(&r.elements, out.size());
                out.push_back({ &r.elements, pass >= 0 ? r.value : std::numeric_limits<int>::min(),
                                pass <= 0 ? r.value : std::numeric_limits<int>::max(), r.id });
                continue;
            }
            merged& m = out[it->second];
            if (pass == 0) {
                if (m.min <= r.value && r.value <= m.max)
                    m.min = m.max = r.value;
                else
                    return "multiple_constraint: an equality contradicts an earlier one over the same elements";
            } else if (pass < 0)
                m.max = std::min(m.max, r.value);
            else
                m.min = std::max(m.min, r.value);
        }
    }
    if (!by_type) // the type orders keep the rows as the passes met them
        std::sort(out.begin(), out.end(), [](const merged& a, const merged& b) { return a.id < b.id; });

    // p1-p4, make_special_merged_constraints (:219-285): rows by the smallest
    // cost / degree ratio of their variables (p1 descending, the others
    // ascending), elements by descending ratio.  std::sort with ties: same
    // library, same sequence of comparisons, same permutation as the reference.
    std::vector<double> ratio;
    if (special) {
        std::vector<int> degree(p.vars.size(), 0);
        for (const merged& m : out)
            for (const auto& e : *m.elements)
                ++degree[e.second];
        ratio.assign(p.vars.size(), 0.0);
        for (const auto& o : p.objective)
            ratio[o.second] = o.first / static_cast<double>(degree[o.second]);
        std::vector<std::pair<int, double>> row_ratio(out.size());
        for (size_t i = 0; i != out.size(); ++i) {
            const auto& el = *out[i].elements;
            row_ratio[i] = { static_cast<int>(i), ratio[el[0].second] };
            for (size_t j = 1; j != el.size(); ++j)
                row_ratio[i].second = std::min(row_ratio[i].second, ratio[el[j].second]);
        }
        if (pre_order == 8)
            std::sort(row_ratio.begin(), row_ratio.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
        else
            std::sort(row_ratio.begin(), row_ratio.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
        for (size_t i = 0; i != out.size(); ++i)
            out[i].id = row_ratio[i].first;
        std::sort(out.begin(), out.end(), [](const merged& a, const merged& b) { return a.id < b.id; });
    }

    p.row_ptr.assign(1, 0);
    p.col.clear(), p.factor.clear(), p.cmin.clear(), p.cmax.clear();
    for (const merged& m : out) {
        std::vector<std::pair<int, int>> el(*m.elements);
        if (special)
            std::sort(el.begin(), el.end(), [&ratio](const auto& a, const auto& b) { return ratio[a.second] > ratio[b.second]; });
        else if (pre_order >= 1) // improve_memory_usage: elements by variable index
            std::sort(el.begin(), el.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
        for (const auto& e : el) {
            p.factor.push_back(e.first);
            p.col.push_back(e.second);
        }
        p.row_ptr.push_back(static_cast<int32_t>(p.col.size()));
        p.cmin.push_back(m.min);
        p.cmax.push_back(m.max);
    }
    return "";
}

int
lp_build(std::string text, int pre_order, bxg_lp** out)
{
    std::unique_ptr<bxg_lp> p(new (std::nothrow) bxg_lp);
    if (!p)
        return fail(BXG_E_MEMORY, "out of host memory");
    std::string err;
    try {
        err = lp_parse(std::move(text), *p);
        if (err.empty())
            err = lp_merge(*p, pre_order);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "out of host memory");
    }
    if (!err.empty())
        return fail(BXG_E_INVALID, "LP text: " + err);
    *out = p.release();
    return BXG_OK;
}

} // namespace

extern "C" {

int
bxg_lp_parse(const char* text, size_t length, int pre_order, bxg_lp** out)
{
    if (!text || !out)
        return fail(BXG_E_INVALID, "bxg_lp_parse: bad argument");
    // no exception crosses the C boundary
    try {
        return lp_build(std::string(text, length), pre_order, out);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "LP text: out of host memory");
    } catch (const std::exception& e) {
        return fail(BXG_E_INVALID, std::string("LP text: ") + e.what());
    } catch (...) {
        return fail(BXG_E_INVALID, "LP text: unexpected failure");
    }
}

int
bxg_lp_read(const char* path, int pre_order, bxg_lp** out)
{
    if (!path || !out)
        return fail(BXG_E_INVALID, "bxg_lp_read: bad argument");
    struct closer
    {
        void operator()(std::FILE* f) const { std::fclose(f); }
    };
    try {
        std::unique_ptr<std::FILE, closer> f(std::fopen(path, "rb"));
        if (!f)
            return fail(BXG_E_INVALID, std::string("LP text: file_not_found: ") + path);
        struct stat st;
        if (fstat(fileno(f.get()), &st) != 0 || S_ISDIR(st.st_mode))
            return fail(BXG_E_INVALID, std::string("LP text: not a readable file: ") + path);
        std::string text;
        if (S_ISREG(st.st_mode) && st.st_size > 0) { // one read when the size is known
            text.resize(static_cast<size_t>(st.st_size));
            text.resize(std::fread(&text[0], 1, text.size(), f.get()));
        }
        char chunk[1 << 16];
        for (size_t got; (got = std::fread(chunk, 1, sizeof chunk, f.get())) > 0;) // pipes, files that grew
            text.append(chunk, got);
        return lp_build(std::move(text), pre_order, out);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "LP text: out of host memory");
    } catch (const std::exception& e) {
        return fail(BXG_E_INVALID, std::string("LP text: ") + e.what());
    } catch (...) {
        return fail(BXG_E_INVALID, "LP text: unexpected failure");
    }
}

void
bxg_lp_destroy(bxg_lp* p)
{
    delete p;
}

int
bxg_lp_info(const bxg_lp* p, int* m, int* n, int* nnz, int* problem_type, int* maximize,
            double* cost_constant, int* n_obj)
{
    if (!p)
        return fail(BXG_E_INVALID, "null LP");
    if (m)
        *m = static_cast<int>(p->cmin.size());
    if (n)
        *n = static_cast<int>(p->vars.size());
    if (nnz)
        *nnz = static_cast<int>(p->col.size());
    if (problem_type)
        *problem_type = p->problem_type;
    if (maximize)
        *maximize = p->maximize;
    if (cost_constant)
        *cost_constant = p->constant;
    if (n_obj)
        *n_obj = static_cast<int>(p->objective.size());
    return BXG_OK;
}

int
bxg_lp_rows(const bxg_lp* p, int32_t* row_ptr, int32_t* col, int32_t* factor, int32_t* cmin, int32_t* cmax)
{
    if (!p || !row_ptr || !col || !factor || !cmin || !cmax)
        return fail(BXG_E_INVALID, "bxg_lp_rows: bad argument");
    std::copy(p->row_ptr.begin(), p->row_ptr.end(), row_ptr);
    std::copy(p->col.begin(), p->col.end(), col);
    std::copy(p->factor.begin(), p->factor.end(), factor);
    std::copy(p->cmin.begin(), p->cmin.end(), cmin);
    std::copy(p->cmax.begin(), p->cmax.end(), cmax);
    return BXG_OK;
}

int
bxg_lp_objective(const bxg_lp* p, int32_t* obj_var, double* obj_factor)
{
    if (!p || !obj_var || !obj_factor)
        return fail(BXG_E_INVALID, "bxg_lp_objective: bad argument");
    for (size_t i = 0; i != p->objective.size(); ++i) {
        obj_factor[i] = p->objective[i].first;
        obj_var[i] = p->objective[i].second;
    }
    return BXG_OK;
}

const char*
bxg_lp_variable(const bxg_lp* p, int i, int* min, int* max, int* type)
{
    if (!p || i < 0 || i >= static_cast<int>(p->vars.size()))
        return nullptr;
    if (min)
        *min = p->vars[i].min;
    if (max)
        *max = p->vars[i].max;
    if (type)
        *type = p->vars[i].type;
    return p->vars[i].name.c_str();
}

int
bxg_lp_variables(const bxg_lp* p, int32_t* min, int32_t* max, int32_t* type)
{
    if (!p || !min || !max || !type)
        return fail(BXG_E_INVALID, "bxg_lp_variables: bad argument");
    for (size_t i = 0; i != p->vars.size(); ++i) {
        min[i] = p->vars[i].min;
        max[i] = p->vars[i].max;
        type[i] = p->vars[i].type;
    }
    return BXG_OK;
}

int
bxg_lp_names(const bxg_lp* p, char* buffer, size_t capacity, size_t* needed)
{
    if (!p || !needed)
        return fail(BXG_E_INVALID, "bxg_lp_names: bad argument");
    size_t total = 0;
    for (const auto& v : p->vars)
        total += v.name.size() + 1;
    *needed = total;
    if (!buffer)
        return BXG_OK;
    if (capacity < total)
        return fail(BXG_E_INVALID, "bxg_lp_names: buffer too small");
    for (const auto& v : p->vars) {
        std::memcpy(buffer, v.name.c_str(), v.name.size() + 1);
        buffer += v.name.size() + 1;
    }
    return BXG_OK;
}

} // extern "C"

#endif
