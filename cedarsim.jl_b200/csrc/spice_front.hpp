// spice_front.hpp -- native netlist front end of the engine: a SPICE- / Spectre-subset reader and
// flattener in C++ behind the C ABI (cb_netlist_*, include/cedarb200.h), so that a binding without a front end of its
// own goes from deck text + sweep values to a compiled circuit and its parameter matrix without the Python host.
//
// What it restates (reference files under /root/reference, the same rules as the Python front end netlist.py / expr.py,
// against which tests/test_native_front.py compares it card by card):
//   * lexical level: title line, `*` comment lines, trailing `;` / `$` comments, `+` continuations, case-insensitive
//     names (SPICE lexer, SpectreNetlistParser.jl/src/spice);
//   * numbers with magnitude suffixes t g meg k m mil u n p f a, trailing unit letters ignored, `1Amp` = 1
//     (src/spectre.jl:383-455); the suffix is folded into the decimal exponent before the conversion, so 0.22u == 0.22e-6
//     bit for bit (test/basic.jl:609-638);
//   * parameter expressions ('...', {...}): C-like precedence, `**` / `^` power, ternary, comparison and logical operators,
//     the functions of src/spectre_env.jl (agauss / gauss return the nominal value: the reference disables the rng);
//   * `.param`, `.subckt` / `.ends` (nested definitions, `params:` defaults, instance overrides that may refer to one
//     another and to the caller's scope), dynamic scoping to the enclosing scope (src/spectre.jl:494-512, test/params.jl),
//     `m=` multiplicities nesting multiplicatively (src/simulate_ir.jl:43-75), `.option` / `.temp` (`temper`), `.include`;
//   * devices R C L V I E G and X; sources DC / AC / PULSE / PWL / SIN (src/spectre_env.jl:15-77, 144-198);
//   * sweep variables: top-level parameters, `x1.par` instance parameters, `r1.r` / `c1.c` / `l1.l` / `v1.dc` / `e1.gain`
//     device values, `temp`; every value that differs between sweep points becomes one column of params[P][B]
//     (what ParamSim makes a runtime parameter, src/circuitodesystem.jl:66-97), a NaN entry keeps the default
//     (`nothing` in a SerialSweep, src/sweeps.jl:18-21).
//   * `.if` / `.elseif` / `.else` / `.endif` on constant top-level parameters, `.lib` sections (`.lib name` ... `.endl`,
//     `.lib "file" name`, also from the file itself, test/basic.jl:312-336), `.model` cards of resistors
//     (r = rsh (l - short) / (w - narrow), src/simpledevices.jl:62-70).
//   * the Spectre-language subset of spectre.py (parse_spectre_into below; cb_netlist_flatten_spectre).
// Not here (the decks that need them go through the Python host, which also owns the Verilog-A compiler): behavioural
// sources, MOSFET / Verilog-A instances (`.hdl`, `ahdl_include`).  They are refused with a message, never skipped.
//
// Host-only code; depends on include/cedarb200.h alone.
#pragma once

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cedarb200.h"

namespace sf {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// A value per sweep point: size 1 = the same for every point
struct Val {
    std::vector<double> v;
    Val() : v(1, 0.0) {}
    Val(double x) : v(1, x) {}
    bool uniform() const { return v.size() == 1; }
    double at(size_t i) const { return v.size() == 1 ? v[0] : v[i]; }
};

template <class F> static Val map1(const Val& a, F f) {
    Val r; r.v.resize(a.v.size());
    for (size_t i = 0; i < a.v.size(); i++) r.v[i] = f(a.v[i]);
    return r;
}
template <class F> static Val map2(const Val& a, const Val& b, F f) {
    const size_t n = std::max(a.v.size(), b.v.size());
    if (a.v.size() != 1 && b.v.size() != 1 && a.v.size() != b.v.size()) throw Error("sweep columns of different lengths");
    Val r; r.v.resize(n);
    for (size_t i = 0; i < n; i++) r.v[i] = f(a.at(i), b.at(i));
    return r;
}

static std::string lower(std::string s) {
    for (char& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// ---- numbers ---------------------------------------------------------------------------------------------------------
// returns the number of characters consumed (0 = not a number)
static size_t scan_number(const std::string& s, size_t pos, double& out) {
    size_t i = pos;
    const size_t n = s.size();
    std::string mant;
    if (i < n && (s[i] == '+' || s[i] == '-')) mant += s[i++];
    size_t digits = 0;
    while (i < n && std::isdigit((unsigned char)s[i])) { mant += s[i++]; digits++; }
    if (i < n && s[i] == '.') {
        mant += s[i++];
        while (i < n && std::isdigit((unsigned char)s[i])) { mant += s[i++]; digits++; }
    }
    if (!digits) return 0;
    long exp10 = 0;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        size_t j = i + 1;
        std::string ex;
        if (j < n && (s[j] == '+' || s[j] == '-')) ex += s[j++];
        size_t ed = 0;
        while (j < n && std::isdigit((unsigned char)s[j])) { ex += s[j++]; ed++; }
        if (ed) { exp10 = std::strtol(ex.c_str(), nullptr, 10); i = j; }
    }
    // magnitude suffix + unit letters
    size_t j = i;
    std::string rest;
    while (j < n && (std::isalpha((unsigned char)s[j]) || s[j] == '_')) rest += (char)std::tolower((unsigned char)s[j++]);
    size_t k = 0;
    while (k < rest.size() && rest[k] == '_') k++;
    rest = rest.substr(k);
    double scale = 1.0;
    if (rest.compare(0, 3, "mil") == 0) scale = 25.4e-6;
    else if (rest.compare(0, 3, "meg") == 0) exp10 += 6;
    else if (!rest.empty() && rest.compare(0, 2, "am") != 0) {   // `1Amp` is one ampere, not one atto-"mp"
        switch (rest[0]) {
            case 't': exp10 += 12; break; case 'g': exp10 += 9; break; case 'k': exp10 += 3; break;
            case 'm': exp10 -= 3; break; case 'u': exp10 -= 6; break; case 'n': exp10 -= 9; break;
            case 'p': exp10 -= 12; break; case 'f': exp10 -= 15; break; case 'a': exp10 -= 18; break;
            default: break;
        }
    }
    out = std::strtod((mant + "e" + std::to_string(exp10)).c_str(), nullptr) * scale;
    return j - pos;
}

static double parse_number(const std::string& tok) {
    std::string s = lower(tok);
    while (!s.empty() && (s.back() == ',' || std::isspace((unsigned char)s.back()))) s.pop_back();
    size_t b = 0;
    while (b < s.size() && std::isspace((unsigned char)s[b])) b++;
    double v = 0;
    const size_t n = scan_number(s, b, v);
    if (!n || b + n != s.size()) throw Error("not a number: '" + tok + "'");
    return v;
}

// ---- expressions -----------------------------------------------------------------------------------------------------
struct Expr;
using ExprP = std::shared_ptr<Expr>;
struct Expr {
    enum Kind { NUM, VAR, NEG, NOT, INV, BIN, COND, CALL } kind = NUM;
    double num = 0;
    std::string name;   // VAR / CALL name, BIN operator
    std::vector<ExprP> a;
};

struct Tok { int kind; double num; std::string s; };   // 0 num, 1 id, 2 op, 3 eof

static std::vector<Tok> tokenize_expr(const std::string& text) {
    std::vector<Tok> out;
    size_t i = 0;
    const size_t n = text.size();
    static const char* ops2[] = {"**", "==", "!=", "<=", ">=", "&&", "||", "~^", "^~", "<<", ">>"};
    while (i < n) {
        if (std::isspace((unsigned char)text[i])) { i++; continue; }
        const char c = text[i];
        if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < n && std::isdigit((unsigned char)text[i + 1]))) {
            double v;
            const size_t k = scan_number(text, i, v);
            if (!k) throw Error("bad number in expression '" + text + "'");
            out.push_back({0, v, ""});
            i += k;
            continue;
        }
        if (std::isalpha((unsigned char)c) || c == '_' || c == '$') {
            size_t j = i + 1;
            while (j < n && (std::isalnum((unsigned char)text[j]) || text[j] == '_' || text[j] == '.' || text[j] == '$')) j++;
            out.push_back({1, 0, lower(text.substr(i, j - i))});
            i = j;
            continue;
        }
        bool two = false;
        if (i + 1 < n)
            for (const char* o : ops2)
                if (text[i] == o[0] && text[i + 1] == o[1]) { out.push_back({2, 0, o}); i += 2; two = true; break; }
        if (two) continue;
        if (std::strchr("-+*/^(),<>?:!&|~", c)) { out.push_back({2, 0, std::string(1, c)}); i++; continue; }
        throw Error("bad expression near '" + text.substr(i, 20) + "'");
    }
    out.push_back({3, 0, ""});
    return out;
}

class ExprParser {
    std::vector<Tok> t;
    size_t i = 0;
    const Tok& peek() const { return t[i]; }
    bool is_op(const char* o) const { return t[i].kind == 2 && t[i].s == o; }
    void eat_op(const char* o) {
        if (!is_op(o)) throw Error(std::string("expected '") + o + "' in expression");
        i++;
    }
    static ExprP mk(Expr::Kind k) { auto e = std::make_shared<Expr>(); e->kind = k; return e; }
    ExprP ternary() {
        ExprP c = binary(0);
        if (is_op("?")) {
            i++;
            ExprP a = ternary();
            eat_op(":");
            ExprP b = ternary();
            ExprP e = mk(Expr::COND);
            e->a = {c, a, b};
            return e;
        }
        return c;
    }
    ExprP binary(int lvl) {
        static const std::vector<std::vector<std::string>> L = {{"||"}, {"&&"}, {"|"}, {"~^", "^~"}, {"&"}, {"==", "!="},
                                                                 {"<", "<=", ">", ">="}, {"<<", ">>"}, {"+", "-"}, {"*", "/"}};
        if (lvl == (int)L.size()) return unary();
        ExprP lhs = binary(lvl + 1);
        for (;;) {
            bool hit = false;
            if (peek().kind == 2)
                for (const std::string& o : L[lvl]) if (peek().s == o) hit = true;
            if (!hit) break;
            ExprP e = mk(Expr::BIN);
            e->name = peek().s;
            i++;
            e->a = {lhs, binary(lvl + 1)};
            lhs = e;
        }
        return lhs;
    }
    ExprP unary() {
        if (is_op("-")) { i++; ExprP e = mk(Expr::NEG); e->a = {unary()}; return e; }
        if (is_op("+")) { i++; return unary(); }
        if (is_op("!")) { i++; ExprP e = mk(Expr::NOT); e->a = {unary()}; return e; }
        if (is_op("~")) { i++; ExprP e = mk(Expr::INV); e->a = {unary()}; return e; }
        return power();
    }
    ExprP power() {
        ExprP base = atom();
        if (is_op("**") || is_op("^")) {
            i++;
            ExprP e = mk(Expr::BIN);
            e->name = "**";
            e->a = {base, unary()};
            return e;
        }
        return base;
    }
    ExprP atom() {
        const Tok tk = peek();
        if (tk.kind == 0) { i++; ExprP e = mk(Expr::NUM); e->num = tk.num; return e; }
        if (tk.kind == 1) {
            i++;
            if (is_op("(")) {
                i++;
                ExprP e = mk(Expr::CALL);
                e->name = tk.s;
                if (!is_op(")")) {
                    e->a.push_back(ternary());
                    while (is_op(",")) { i++; e->a.push_back(ternary()); }
                }
                eat_op(")");
                return e;
            }
            ExprP e = mk(Expr::VAR);
            e->name = tk.s;
            return e;
        }
        if (is_op("(")) { i++; ExprP e = ternary(); eat_op(")"); return e; }
        throw Error("unexpected token '" + tk.s + "' in expression");
    }

public:
    static ExprP parse(std::string text) {
        size_t b = 0, e = text.size();
        while (b < e && std::isspace((unsigned char)text[b])) b++;
        while (e > b && std::isspace((unsigned char)text[e - 1])) e--;
        text = text.substr(b, e - b);
        if (text.size() >= 2 && std::strchr("'{\"", text.front()) && std::strchr("'}\"", text.back())) text = text.substr(1, text.size() - 2);
        ExprParser p;
        p.t = tokenize_expr(text);
        ExprP r = p.ternary();
        if (p.peek().kind != 3) throw Error("trailing input in expression '" + text + "'");
        return r;
    }
};

static bool const_value(const std::string& name, double& out) {
    static const std::map<std::string, double> C = {{"pi", M_PI}, {"e", M_E}, {"true", 1.0}, {"false", 0.0}, {"m_1_pi", 1.0 / M_PI},
                                                    {"m_pi", M_PI}, {"m_e", M_E}, {"m_two_pi", 2.0 * M_PI}, {"m_sqrt2", std::sqrt(2.0)}};
    auto it = C.find(name);
    if (it == C.end()) return false;
    out = it->second;
    return true;
}

struct Scope;
static Val eval_expr(const ExprP& e, Scope& sc);

// Parameter scope with lazy, memoised evaluation and dynamic scoping to the parent (netlist.py _Scope)
struct Scope {
    std::map<std::string, std::string> exprs;
    Scope* parent = nullptr;
    std::map<std::string, Val> values;
    std::set<std::string> busy;
    bool lookup(const std::string& name, Val& out) {
        auto v = values.find(name);
        if (v != values.end()) { out = v->second; return true; }
        auto e = exprs.find(name);
        if (e != exprs.end()) {
            if (busy.count(name)) {
                // `.subckt inner a b foo=foo+2000`: inside its own default a name means the enclosing scope's value
                if (parent) return parent->lookup(name, out);
                throw Error("circular parameter definition for '" + name + "'");
            }
            busy.insert(name);
            Val r = eval_expr(ExprParser::parse(e->second), *this);
            busy.erase(name);
            values[name] = r;
            out = r;
            return true;
        }
        if (parent) return parent->lookup(name, out);
        return false;
    }
    Val eval(const std::string& text) { return eval_expr(ExprParser::parse(text), *this); }
};

static long long to_int(double x) { return (long long)x; }

static Val eval_expr(const ExprP& e, Scope& sc) {
    switch (e->kind) {
        case Expr::NUM: return Val(e->num);
        case Expr::VAR: {
            Val v;
            if (sc.lookup(e->name, v)) return v;
            double c;
            if (const_value(e->name, c)) return Val(c);
            throw Error("undefined parameter '" + e->name + "'");
        }
        case Expr::NEG: return map1(eval_expr(e->a[0], sc), [](double x) { return -x; });
        case Expr::NOT: return map1(eval_expr(e->a[0], sc), [](double x) { return x != 0 ? 0.0 : 1.0; });
        case Expr::INV: return map1(eval_expr(e->a[0], sc), [](double x) { return (double)(~to_int(x)); });
        case Expr::COND: {
            const Val c = eval_expr(e->a[0], sc), a = eval_expr(e->a[1], sc), b = eval_expr(e->a[2], sc);
            const size_t n = std::max(c.v.size(), std::max(a.v.size(), b.v.size()));
            Val r; r.v.resize(n);
            for (size_t i = 0; i < n; i++) r.v[i] = c.at(i) != 0 ? a.at(i) : b.at(i);
            return r;
        }
        case Expr::BIN: {
            const Val a = eval_expr(e->a[0], sc), b = eval_expr(e->a[1], sc);
            const std::string& op = e->name;
            if (op == "+") return map2(a, b, [](double x, double y) { return x + y; });
            if (op == "-") return map2(a, b, [](double x, double y) { return x - y; });
            if (op == "*") return map2(a, b, [](double x, double y) { return x * y; });
            if (op == "/") return map2(a, b, [](double x, double y) { return x / y; });
            if (op == "**") return map2(a, b, [](double x, double y) { return std::pow(x, y); });
            if (op == "==") return map2(a, b, [](double x, double y) { return x == y ? 1.0 : 0.0; });
            if (op == "!=") return map2(a, b, [](double x, double y) { return x != y ? 1.0 : 0.0; });
            if (op == "<") return map2(a, b, [](double x, double y) { return x < y ? 1.0 : 0.0; });
            if (op == "<=") return map2(a, b, [](double x, double y) { return x <= y ? 1.0 : 0.0; });
            if (op == ">") return map2(a, b, [](double x, double y) { return x > y ? 1.0 : 0.0; });
            if (op == ">=") return map2(a, b, [](double x, double y) { return x >= y ? 1.0 : 0.0; });
            if (op == "&&") return map2(a, b, [](double x, double y) { return (x != 0 && y != 0) ? 1.0 : 0.0; });
            if (op == "||") return map2(a, b, [](double x, double y) { return (x != 0 || y != 0) ? 1.0 : 0.0; });
            if (op == "&") return map2(a, b, [](double x, double y) { return (double)(to_int(x) & to_int(y)); });
            if (op == "|") return map2(a, b, [](double x, double y) { return (double)(to_int(x) | to_int(y)); });
            if (op == "~^" || op == "^~") return map2(a, b, [](double x, double y) { return (double)(~(to_int(x) ^ to_int(y))); });
            if (op == "<<") return map2(a, b, [](double x, double y) { return (double)(to_int(x) << to_int(y)); });
            if (op == ">>") return map2(a, b, [](double x, double y) { return (double)(to_int(x) >> to_int(y)); });
            throw Error("unknown operator " + op);
        }
        case Expr::CALL: {
            std::vector<Val> args;
            for (const ExprP& x : e->a) args.push_back(eval_expr(x, sc));
            const std::string& f = e->name;
            auto need = [&](size_t n) { if (args.size() != n) throw Error("wrong number of arguments to " + f + "()"); };
            typedef double (*F1)(double);
            static const std::map<std::string, F1> one = {
                {"sqrt", [](double x) { return std::sqrt(x); }}, {"exp", [](double x) { return std::exp(x); }},
                {"ln", [](double x) { return std::log(x); }}, {"log", [](double x) { return std::log(x); }},
                {"log10", [](double x) { return std::log10(x); }}, {"abs", [](double x) { return std::fabs(x); }},
                {"sin", [](double x) { return std::sin(x); }}, {"cos", [](double x) { return std::cos(x); }},
                {"tan", [](double x) { return std::tan(x); }}, {"atan", [](double x) { return std::atan(x); }},
                {"arctan", [](double x) { return std::atan(x); }}, {"sinh", [](double x) { return std::sinh(x); }},
                {"cosh", [](double x) { return std::cosh(x); }}, {"tanh", [](double x) { return std::tanh(x); }},
                {"asinh", [](double x) { return std::asinh(x); }}, {"acosh", [](double x) { return std::acosh(x); }},
                {"atanh", [](double x) { return std::atanh(x); }}, {"floor", [](double x) { return std::floor(x); }},
                {"ceil", [](double x) { return std::ceil(x); }}, {"int", [](double x) { return std::trunc(x); }},
                {"nint", [](double x) { return std::nearbyint(x); }}};
            auto it = one.find(f);
            if (it != one.end()) { need(1); F1 fn = it->second; return map1(args[0], [fn](double x) { return fn(x); }); }
            if (f == "min") { need(2); return map2(args[0], args[1], [](double x, double y) { return std::fmin(x, y); }); }
            if (f == "max") { need(2); return map2(args[0], args[1], [](double x, double y) { return std::fmax(x, y); }); }
            if (f == "pow" || f == "pwr") { need(2); return map2(args[0], args[1], [](double x, double y) { return std::pow(x, y); }); }
            if (f == "agauss" || f == "gauss") { need(3); return args[0]; }   // rng disabled in the reference (src/spectre_env.jl:178-187)
            throw Error("unknown function '" + f + "'");
        }
    }
    throw Error("bad expression node");
}

// ---- lexical level of the deck ---------------------------------------------------------------------------------------
static std::vector<std::string> logical_lines(const std::string& text, bool first_is_title) {
    std::vector<std::string> raw;
    {
        std::string cur;
        for (char c : text) {
            if (c == '\n') { raw.push_back(cur); cur.clear(); }
            else if (c != '\r') cur += c;
        }
        if (!cur.empty()) raw.push_back(cur);
    }
    std::vector<std::string> out;
    for (size_t li = (first_is_title && !raw.empty()) ? 1 : 0; li < raw.size(); li++) {
        std::string s = raw[li];
        size_t b = 0, e = s.size();
        while (b < e && std::isspace((unsigned char)s[b])) b++;
        while (e > b && std::isspace((unsigned char)s[e - 1])) e--;
        s = s.substr(b, e - b);
        if (s.empty() || s[0] == '*') continue;
        // trailing comments: `;` or `$` after white space (or at the start)
        {
            const std::string t = " " + s;
            for (size_t k = 0; k + 1 < t.size(); k++)
                if (std::isspace((unsigned char)t[k]) && (t[k + 1] == ';' || t[k + 1] == '$')) { s = t.substr(1, k > 0 ? k - 1 : 0); break; }
            while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
        }
        if (s.empty()) continue;
        if (s[0] == '+') {
            if (!out.empty()) {
                size_t k = 1;
                while (k < s.size() && std::isspace((unsigned char)s[k])) k++;
                out.back() += " " + s.substr(k);
            }
            continue;
        }
        out.push_back(s);
    }
    return out;
}

// '...' | {...} | "..." | ( | ) | , | = | run of other characters
static std::vector<std::string> tokens(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    const size_t n = line.size();
    while (i < n) {
        const char c = line[i];
        if (std::isspace((unsigned char)c)) { i++; continue; }
        if (c == '\'' || c == '"' || c == '{') {
            const char close = c == '{' ? '}' : c;
            size_t j = line.find(close, i + 1);
            if (j == std::string::npos) throw Error("unterminated quote in '" + line + "'");
            out.push_back(line.substr(i, j - i + 1));
            i = j + 1;
            continue;
        }
        if (c == '(' || c == ')' || c == ',' || c == '=') { out.push_back(std::string(1, c)); i++; continue; }
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)line[j]) && !std::strchr("(),=", line[j])) j++;
        out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

typedef std::vector<std::pair<std::string, std::string>> KV;   // ordered k = v pairs (keys lower case)

static void split_params(const std::vector<std::string>& toks, size_t from, std::vector<std::string>& pos, KV& kv) {
    size_t i = from;
    while (i < toks.size()) {
        if (i + 1 < toks.size() && toks[i + 1] == "=") {
            const std::string key = lower(toks[i]), val = i + 2 < toks.size() ? toks[i + 2] : "";
            bool found = false;
            for (auto& p : kv) if (p.first == key) { p.second = val; found = true; }
            if (!found) kv.push_back({key, val});
            i += 3;
        } else pos.push_back(toks[i++]);
    }
}
static const std::string* kv_get(const KV& kv, const std::string& key) {
    for (auto& p : kv) if (p.first == key) return &p.second;
    return nullptr;
}

struct Source { bool has_dc = false, has_ac = false; std::string dc, ac, tran_kind; std::vector<std::string> tran_args; };

struct Card {
    char kind = 0;
    std::string name;
    std::vector<std::string> nodes;
    std::string model;            // X: subcircuit name
    bool has_value = false;
    std::string value;
    KV params;
    Source src;
};

struct Subckt {
    std::string name;
    std::vector<std::string> ports;
    KV params, local_params;
    std::vector<Card> cards;
    std::map<std::string, std::shared_ptr<Subckt>> subckts;
};

struct Netlist {
    Subckt top;
    KV options;
    std::map<std::string, std::map<std::string, double>> models;   // .model cards: name -> numeric parameters (lower-case keys)
    std::map<std::string, std::string> model_master;               // name -> master (r, c, nmos, ...)
};

static Source parse_source(const std::vector<std::string>& toks, size_t from) {
    Source s;
    std::vector<std::string> flat;
    for (size_t i = from; i < toks.size(); i++) if (toks[i] != "," && toks[i] != "=") flat.push_back(toks[i]);
    size_t i = 0;
    while (i < flat.size()) {
        const std::string t = lower(flat[i]);
        if (t == "dc" && i + 1 < flat.size()) { s.has_dc = true; s.dc = flat[i + 1]; i += 2; }
        else if (t == "ac" && i + 1 < flat.size()) {
            s.has_ac = true; s.ac = flat[i + 1]; i += 2;
            if (i < flat.size() && (std::isdigit((unsigned char)flat[i][0]) || std::strchr("-+.", flat[i][0]))) i++;   // phase
        } else if (t == "pwl" || t == "pulse" || t == "sin") {
            s.tran_kind = t;
            i++;
            if (i < flat.size() && flat[i] == "(") {
                i++;
                while (i < flat.size() && flat[i] != ")") s.tran_args.push_back(flat[i++]);
                i++;
            } else {
                while (i < flat.size() && lower(flat[i]) != "dc" && lower(flat[i]) != "ac") s.tran_args.push_back(flat[i++]);
            }
        } else if (t == "(" || t == ")") i++;
        else {
            if (!s.has_dc && s.tran_kind.empty()) { s.has_dc = true; s.dc = flat[i]; }   // a bare value means DC
            i++;
        }
    }
    return s;
}

static Card parse_card(const std::vector<std::string>& toks) {
    Card c;
    c.name = lower(toks[0]);
    c.kind = c.name[0];
    auto unsupported = [&](const char* what) -> Error {
        return Error(c.name + ": " + what + " are handled by the Python front end (netlist.py), not by the native reader");
    };
    if (std::strchr("rcl", c.kind)) {
        std::vector<std::string> pos;
        split_params(toks, 1, pos, c.params);
        if (pos.size() < 2) throw Error(c.name + ": two nodes expected");
        c.nodes = {lower(pos[0]), lower(pos[1])};
        if (pos.size() > 2) {
            const std::string& v = pos[2];
            if ((std::isalpha((unsigned char)v[0]) || v[0] == '_')) {
                // a bare identifier: a model name (`R1 a b rmod l=2u w=1u`), or a parameter (`R2 vcc 0 res`, test/basic.jl:725-737)
                // -- decided when the circuit is flattened; an optional value may follow the model name
                c.model = lower(v);
                if (pos.size() > 3) { c.has_value = true; c.value = pos[3]; }
            } else { c.has_value = true; c.value = v; }
        }
        return c;
    }
    if (c.kind == 'v' || c.kind == 'i') {
        if (toks.size() < 3) throw Error(c.name + ": two nodes expected");
        c.nodes = {lower(toks[1]), lower(toks[2])};
        c.src = parse_source(toks, 3);
        return c;
    }
    if (c.kind == 'e' || c.kind == 'g') {
        for (size_t i = 3; i < toks.size() && i < 5; i++) {
            const std::string t = lower(toks[i]);
            if (t == "vol" || t == "cur" || t == "value") throw unsupported("behavioural sources");
        }
        std::vector<std::string> flat{toks[0]}, pos;
        for (size_t i = 1; i < toks.size(); i++) if (toks[i] != "(" && toks[i] != ")" && toks[i] != ",") flat.push_back(toks[i]);
        split_params(flat, 1, pos, c.params);
        if (pos.size() < 4) throw Error(c.name + ": four nodes expected");
        for (int k = 0; k < 4; k++) c.nodes.push_back(lower(pos[k]));
        if (pos.size() > 4) { c.has_value = true; c.value = pos[4]; }
        else if (const std::string* g = kv_get(c.params, "gain")) { c.has_value = true; c.value = *g; }
        return c;
    }
    if (c.kind == 'x') {
        std::vector<std::string> pos;
        split_params(toks, 1, pos, c.params);
        if (pos.empty()) throw Error(c.name + ": subcircuit name expected");
        c.model = lower(pos.back());
        for (size_t k = 0; k + 1 < pos.size(); k++) c.nodes.push_back(lower(pos[k]));
        return c;
    }
    if (c.kind == 'b') throw unsupported("behavioural sources");
    if (c.kind == 'm') throw unsupported("MOSFET instances (generated Verilog-A device code)");
    throw Error("unsupported device card '" + toks[0] + "'");
}

// values of the top-level parameters that are constants (for `.if` conditions, evaluated while reading like the reference)
static Val const_env_eval(Netlist& nl, const std::string& text) {
    Scope sc;
    for (auto& p : nl.top.params) {
        try {
            Scope one;
            one.values = sc.values;
            sc.values[p.first] = one.eval(p.second);
        } catch (const Error&) {}
    }
    return sc.eval(text);
}

// lib_section: empty = an ordinary file; else only the `.lib <name>` ... `.endl` section of that name is read
static void parse_into(Netlist& nl, const std::string& text, bool first_is_title, const std::string& base_dir, int depth,
                       const std::string& lib_section = std::string()) {
    if (depth > 16) throw Error(".include nesting too deep");
    std::vector<Subckt*> stack{&nl.top};
    std::vector<char> cond_stack, cond_taken;   // .if nesting: is the current branch live / has any branch been live
    std::string in_lib;                         // name of the `.lib` section being defined here, if any
    bool in_lib_def = false;
    for (const std::string& line : logical_lines(text, first_is_title)) {
        const std::vector<std::string> toks = tokens(line);
        if (toks.empty()) continue;
        const std::string head = lower(toks[0]);
        Subckt* cur = stack.back();
        auto unquote = [](std::string f) {
            if (f.size() >= 2 && std::strchr("'\"", f.front())) f = f.substr(1, f.size() - 2);
            return f;
        };
        if (head == ".lib" && toks.size() == 2) { in_lib = lower(unquote(toks[1])); in_lib_def = true; continue; }   // section definition
        if (head == ".endl") { in_lib_def = false; in_lib.clear(); continue; }
        // a section is inert where it is defined and is read only through `.lib "file" name` (also from the file itself)
        if (lib_section.empty() ? in_lib_def : (!in_lib_def || in_lib != lib_section)) continue;
        if (head == ".if" || head == ".elseif" || head == ".else" || head == ".endif") {
            std::string cond;   // from the raw line: the card tokenizer splits `==`
            {
                size_t k = 0;
                while (k < line.size() && std::isspace((unsigned char)line[k])) k++;
                while (k < line.size() && !std::isspace((unsigned char)line[k]) && line[k] != '(') k++;
                cond = line.substr(k);
            }
            auto truth = [&]() { const Val v = const_env_eval(nl, cond); return v.v[0] != 0.0; };
            if (head == ".if") { cond_stack.push_back(truth()); cond_taken.push_back(cond_stack.back()); }
            else if (cond_stack.empty()) throw Error(head + " without .if");
            else if (head == ".else") { cond_stack.back() = !cond_taken.back(); cond_taken.back() = 1; }
            else if (head == ".elseif") { cond_stack.back() = !cond_taken.back() && truth(); cond_taken.back() = cond_taken.back() || cond_stack.back(); }
            else { cond_stack.pop_back(); cond_taken.pop_back(); }
            continue;
        }
        {
            bool live = true;
            for (char c : cond_stack) live = live && c;
            if (!live) continue;
        }
        if (head[0] == '.') {
            if (head == ".param" || head == ".parameter" || head == ".parameters") {
                std::vector<std::string> pos;
                KV kv;
                split_params(toks, 1, pos, kv);
                KV& dst = cur == &nl.top ? cur->params : cur->local_params;
                for (auto& p : kv) {
                    bool found = false;
                    for (auto& q : dst) if (q.first == p.first) { q.second = p.second; found = true; }
                    if (!found) dst.push_back(p);
                }
            } else if (head == ".subckt") {
                std::vector<std::string> pos;
                KV kv;
                split_params(toks, 1, pos, kv);
                std::vector<std::string> p2;
                for (auto& p : pos) if (lower(p) != "params:") p2.push_back(lower(p));
                if (p2.empty()) throw Error(".subckt without a name");
                auto sub = std::make_shared<Subckt>();
                sub->name = p2[0];
                sub->ports.assign(p2.begin() + 1, p2.end());
                sub->params = kv;
                cur->subckts[sub->name] = sub;
                stack.push_back(sub.get());
            } else if (head == ".ends") {
                if (stack.size() > 1) stack.pop_back();
            } else if (head == ".include" || head == ".inc" || head == ".lib") {
                if (toks.size() < 2) throw Error(head + " without a file name");
                std::string fname = unquote(toks[1]);
                const std::string section = (head == ".lib" && toks.size() > 2) ? lower(unquote(toks[2])) : std::string();
                if (lower(fname).compare(0, 8, "jlpkg://") == 0) throw Error("package include " + fname + " is handled by the Python front end");
                const std::string path = (!fname.empty() && fname[0] == '/') ? fname : (base_dir.empty() ? fname : base_dir + "/" + fname);
                std::ifstream f(path);
                if (!f) throw Error("cannot open include file '" + path + "'");
                std::stringstream ss;
                ss << f.rdbuf();
                const size_t slash = path.find_last_of('/');
                parse_into(nl, ss.str(), false, slash == std::string::npos ? std::string() : path.substr(0, slash), depth + 1, section);
            } else if (head == ".option" || head == ".options") {
                std::vector<std::string> pos;
                split_params(toks, 1, pos, nl.options);
            } else if (head == ".temp") {
                if (toks.size() > 1) {
                    bool found = false;
                    for (auto& q : nl.options) if (q.first == "temp") { q.second = toks[1]; found = true; }
                    if (!found) nl.options.push_back({"temp", toks[1]});
                }
            } else if (head == ".model") {
                // `.model name master (k=v ...)`: numeric parameters are kept (resistor geometry cards); device models that
                // need generated code are refused where an instance uses them
                std::vector<std::string> flat, pos;
                for (size_t q = 1; q < toks.size(); q++) if (toks[q] != "(" && toks[q] != ")" && toks[q] != ",") flat.push_back(toks[q]);
                KV kv;
                split_params(flat, 0, pos, kv);
                if (pos.size() < 2) throw Error(".model needs a name and a master");
                const std::string name = lower(pos[0]);
                nl.model_master[name] = lower(pos[1]);
                std::map<std::string, double>& mp = nl.models[name];
                for (auto& q : kv) {
                    try {
                        Scope empty;
                        const Val v = empty.eval(q.second);
                        mp[q.first] = v.v[0];
                    } catch (const Error&) {}   // not a constant: left to the (absent) parameter scope, like the Python reader's card.exprs
                }
            } else if (head == ".hdl") {
                throw Error(head + " cards are handled by the Python front end (netlist.py), not by the native reader");
            }
            // every other dot-card is ignored, as the reference warns and continues (src/spectre.jl:1520-1522)
            continue;
        }
        cur->cards.push_back(parse_card(toks));
    }
}

// ---- Spectre-language subset (the same rules as spectre.py; reference's Spectre-syntax tests test/basic.jl:168-205,
//      :265-278): instances `name (n1 n2 ...) master k=v ...` of resistor capacitor inductor vsource isource vcvs vccs and of
//      subcircuits, `subckt ... ends` with `parameters`, `model` cards of resistors, `include`, `type=pwl wave=[...]` /
//      `type=sine` / `type=pulse` sources, `\` continuations, `//` and `*` comments.  bsource / ahdl_include are refused.
static std::vector<std::string> spectre_lines(const std::string& text) {
    std::vector<std::string> out;
    std::string cur, raw;
    auto flush_raw = [&]() {
        std::string line;   // strip `// ...` outside quotes
        {
            char quote = 0;
            size_t k = 0;
            for (; k < raw.size(); k++) {
                const char ch = raw[k];
                if (quote) { if (ch == quote) quote = 0; }
                else if (ch == '"' || ch == '\'') quote = ch;
                else if (ch == '/' && k + 1 < raw.size() && raw[k + 1] == '/') break;
            }
            line = raw.substr(0, k);
        }
        while (!line.empty() && std::isspace((unsigned char)line.back())) line.pop_back();
        size_t b = 0;
        while (b < line.size() && std::isspace((unsigned char)line[b])) b++;
        if (cur.empty() && b < line.size() && line[b] == '*') return;
        if (!line.empty() && line.back() == '\\') { cur += line.substr(0, line.size() - 1) + " "; return; }
        cur += line;
        long open = 0;
        for (char ch : cur) open += (ch == '[') - (ch == ']');
        if (open > 0) { cur += " "; return; }   // a vector that runs over the line end
        size_t c0 = 0, c1 = cur.size();
        while (c0 < c1 && std::isspace((unsigned char)cur[c0])) c0++;
        while (c1 > c0 && std::isspace((unsigned char)cur[c1 - 1])) c1--;
        if (c1 > c0) out.push_back(cur.substr(c0, c1 - c0));
        cur.clear();
    };
    for (char c : text) {
        if (c == '\n') { flush_raw(); raw.clear(); }
        else if (c != '\r') raw += c;
    }
    flush_raw();
    {
        size_t c0 = 0, c1 = cur.size();
        while (c0 < c1 && std::isspace((unsigned char)cur[c0])) c0++;
        while (c1 > c0 && std::isspace((unsigned char)cur[c1 - 1])) c1--;
        if (c1 > c0) out.push_back(cur.substr(c0, c1 - c0));
    }
    return out;
}

// [...] | "..." | '...' | ( | ) | = | run of other characters
static std::vector<std::string> spectre_tokens(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    const size_t n = line.size();
    while (i < n) {
        const char c = line[i];
        if (std::isspace((unsigned char)c)) { i++; continue; }
        if (c == '[' || c == '"' || c == '\'') {
            const char close = c == '[' ? ']' : c;
            size_t j = line.find(close, i + 1);
            if (j == std::string::npos) throw Error("unterminated bracket / quote in '" + line + "'");
            out.push_back(line.substr(i, j - i + 1));
            i = j + 1;
            continue;
        }
        if (c == '(' || c == ')' || c == '=') { out.push_back(std::string(1, c)); i++; continue; }
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)line[j]) && !std::strchr("()=[]", line[j])) j++;
        out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

// `a=1 b = x*2 c = p ? 1 : 2` -> (a, '1'), (b, 'x*2'), (c, 'p ? 1 : 2'): a value runs up to the next `name =`
static KV spectre_assignments(const std::string& text) {
    struct Hit { size_t name0, name1, eq1; };
    std::vector<Hit> hits;
    const size_t n = text.size();
    for (size_t i = 0; i < n; i++) {
        if (!(std::isalpha((unsigned char)text[i]) || text[i] == '_')) continue;
        if (i > 0 && (std::isalnum((unsigned char)text[i - 1]) || text[i - 1] == '_' || text[i - 1] == '.' || text[i - 1] == '$')) continue;
        size_t j = i;
        while (j < n && (std::isalnum((unsigned char)text[j]) || text[j] == '_')) j++;
        size_t k = j;
        while (k < n && std::isspace((unsigned char)text[k])) k++;
        if (k < n && text[k] == '=' && !(k + 1 < n && text[k + 1] == '=')) hits.push_back({i, j, k + 1});   // not `==` (and `!=` `<=` `>=` never follow a name directly... they follow an operator character)
        i = j > i ? j - 1 : i;
    }
    KV out;
    for (size_t h = 0; h < hits.size(); h++) {
        const size_t end = h + 1 < hits.size() ? hits[h + 1].name0 : n;
        std::string val = text.substr(hits[h].eq1, end - hits[h].eq1);
        size_t b = 0, e = val.size();
        while (b < e && std::isspace((unsigned char)val[b])) b++;
        while (e > b && std::isspace((unsigned char)val[e - 1])) e--;
        const std::string key = lower(text.substr(hits[h].name0, hits[h].name1 - hits[h].name0));
        bool found = false;
        for (auto& p : out) if (p.first == key) { p.second = val.substr(b, e - b); found = true; }
        if (!found) out.push_back({key, val.substr(b, e - b)});
    }
    return out;
}

static std::string strip_quotes(std::string f) {
    while (!f.empty() && std::strchr("\"'", f.front())) f.erase(f.begin());
    while (!f.empty() && std::strchr("\"'", f.back())) f.pop_back();
    return f;
}

static Source spectre_source(const KV& kv, const std::string& name) {
    Source s;
    auto get = [&](const char* k, const char* d) { const std::string* v = kv_get(kv, k); return v ? *v : std::string(d); };
    if (const std::string* v = kv_get(kv, "dc")) { s.has_dc = true; s.dc = *v; }
    if (const std::string* v = kv_get(kv, "mag")) { s.has_ac = true; s.ac = *v; }
    const std::string typ = lower(strip_quotes(get("type", "dc")));
    if (typ == "pwl") {
        std::string w = get("wave", "[]");
        std::string flat;
        for (char c : w) flat += (c == '[' || c == ']' || c == ',') ? ' ' : c;
        std::stringstream ss(flat);
        std::string t;
        s.tran_kind = "pwl";
        while (ss >> t) s.tran_args.push_back(t);
    } else if (typ == "sine" || typ == "sin") {   // SIN(vo va freq td theta phase), src/spectre_env.jl:169-176
        s.tran_kind = "sin";
        s.tran_args = {kv_get(kv, "sinedc") ? get("sinedc", "0") : get("dc", "0"), get("ampl", "0"), get("freq", "0"), get("delay", "0"),
                       get("damp", "0"), get("sinephase", "0")};
    } else if (typ == "pulse") {                  // PULSE(v1 v2 td tr tf pw per), src/spectre_env.jl:153-166
        s.tran_kind = "pulse";
        s.tran_args = {get("val0", "0"), get("val1", "0"), get("delay", "0"), get("rise", "0"), get("fall", "0"), get("width", "0"),
                       get("period", "0")};
    } else if (typ != "dc") throw Error(name + ": source type '" + typ + "' is outside the Spectre subset");
    return s;
}

static void parse_spectre_into(Netlist& nl, const std::string& text, const std::string& base_dir, int depth) {
    if (depth > 16) throw Error("include nesting too deep");
    static const std::map<std::string, char> prims = {{"resistor", 'r'}, {"capacitor", 'c'}, {"inductor", 'l'}, {"vsource", 'v'},
                                                      {"isource", 'i'}, {"vcvs", 'e'}, {"vccs", 'g'}};
    std::vector<Subckt*> stack{&nl.top};
    for (const std::string& line : spectre_lines(text)) {
        const std::vector<std::string> toks = spectre_tokens(line);
        if (toks.empty()) continue;
        const std::string head = lower(toks[0]);
        Subckt* cur = stack.back();
        if (head == "simulator") {
            if (lower(line).find("spice") != std::string::npos)
                throw Error("`simulator lang=spice` sections are outside the Spectre subset; use the SPICE reader");
            continue;
        }
        if (head == "parameters") {
            KV& dst = cur == &nl.top ? cur->params : cur->local_params;
            for (auto& p : spectre_assignments(line.substr(toks[0].size()))) {
                bool found = false;
                for (auto& q : dst) if (q.first == p.first) { q.second = p.second; found = true; }
                if (!found) dst.push_back(p);
            }
            continue;
        }
        if (head == "subckt" || head == "inline") {
            std::vector<std::string> t;
            for (size_t k = 1; k < toks.size(); k++) if (toks[k] != "(" && toks[k] != ")" && lower(toks[k]) != "subckt") t.push_back(lower(toks[k]));
            if (t.empty()) throw Error("subckt without a name");
            auto sub = std::make_shared<Subckt>();
            sub->name = t[0];
            sub->ports.assign(t.begin() + 1, t.end());
            cur->subckts[sub->name] = sub;
            stack.push_back(sub.get());
            continue;
        }
        if (head == "ends") { if (stack.size() > 1) stack.pop_back(); continue; }
        if (head == "model") {
            if (toks.size() < 3) throw Error("model needs a name and a master");
            const std::string name = lower(toks[1]);
            nl.model_master[name] = lower(toks[2]);
            std::map<std::string, double>& mp = nl.models[name];
            for (auto& q : spectre_assignments(line)) {
                try { Scope empty; mp[q.first] = empty.eval(q.second).v[0]; } catch (const Error&) {}
            }
            continue;
        }
        if (head == "ahdl_include") throw Error("ahdl_include (Verilog-A) is handled by the Python front end");
        if (head == "include") {
            if (toks.size() < 2) throw Error("include without a file name");
            const std::string fname = strip_quotes(toks[1]);
            if (lower(fname).compare(0, 8, "jlpkg://") == 0 || (fname.size() > 3 && lower(fname).substr(fname.size() - 3) == ".va"))
                throw Error("include " + fname + " is handled by the Python front end");
            const std::string path = (!fname.empty() && fname[0] == '/') ? fname : (base_dir.empty() ? fname : base_dir + "/" + fname);
            std::ifstream f(path);
            if (!f) throw Error("cannot open include file '" + path + "'");
            std::stringstream ss;
            ss << f.rdbuf();
            const size_t slash = path.find_last_of('/');
            const std::string dir = slash == std::string::npos ? std::string() : path.substr(0, slash);
            const std::string low = lower(ss.str());
            const bool is_spectre = low.find("simulator lang=spectre") != std::string::npos || (path.size() > 4 && lower(path).substr(path.size() - 4) == ".scs");
            if (is_spectre) parse_spectre_into(nl, ss.str(), dir, depth + 1);
            else parse_into(nl, ss.str(), false, dir, depth + 1, std::string());
            continue;
        }
        if (head == "global" || head == "save" || head == "ic" || head == "nodeset" || head == "options") continue;
        if (toks.size() > 1) {
            const std::string t1 = lower(toks[1]);
            if (t1 == "options" || t1 == "tran" || t1 == "dc" || t1 == "ac" || t1 == "noise" || t1 == "info") continue;   // analyses and controls
        }
        // ---- instance: name (nodes) master k=v ...   (parentheses optional)
        Card c;
        c.name = head;
        std::vector<std::string> nodes, pos;
        KV kv;
        std::string master;
        size_t from = 1;
        if (toks.size() > 1 && toks[1] == "(") {
            size_t close = 2;
            while (close < toks.size() && toks[close] != ")") nodes.push_back(toks[close++]);
            from = close + 1;
            split_params(toks, from, pos, kv);
            master = pos.empty() ? "" : pos[0];
        } else {
            split_params(toks, 1, pos, kv);
            if (!pos.empty()) { master = pos.back(); nodes.assign(pos.begin(), pos.end() - 1); }
        }
        if (master.empty()) throw Error(c.name + ": no master / subcircuit name");
        {   // values may be whole expressions (`r=(p1+p2)/p3`): re-read the assignments from the raw text behind the master
            const size_t at = line.find(master, toks[0].size());
            if (at != std::string::npos) {
                const KV kv2 = spectre_assignments(line.substr(at + master.size()));
                if (!kv2.empty()) kv = kv2;
            }
        }
        for (auto& nd : nodes) nd = lower(nd);
        const std::string m = lower(master);
        auto pit = prims.find(m);
        if (pit != prims.end()) {
            c.kind = pit->second;
            const size_t want = (c.kind == 'e' || c.kind == 'g') ? 4 : 2;
            if (nodes.size() < want) throw Error(c.name + ": " + std::to_string(want) + " nodes expected");
            c.nodes.assign(nodes.begin(), nodes.begin() + want);
            if (c.kind == 'v' || c.kind == 'i') c.src = spectre_source(kv, c.name);
            else if (c.kind == 'e' || c.kind == 'g') {
                const std::string* g = c.kind == 'e' ? kv_get(kv, "gain") : (kv_get(kv, "gm") ? kv_get(kv, "gm") : kv_get(kv, "gain"));
                if (g) { c.has_value = true; c.value = *g; }
                c.params = kv;
            } else c.params = kv;
        } else if (m == "bsource") {
            throw Error(c.name + ": behavioural sources are handled by the Python front end (netlist.py), not by the native reader");
        } else if (nl.models.count(m)) {
            const std::string& master2 = nl.model_master[m];
            if (master2 == "resistor" || master2 == "r" || master2 == "res") {
                if (nodes.size() < 2) throw Error(c.name + ": two nodes expected");
                c.kind = 'r'; c.nodes.assign(nodes.begin(), nodes.begin() + 2); c.model = m; c.params = kv;
            } else throw Error(c.name + ": instances of model '" + m + "' (" + master2 + ") need generated device code: handled by the Python front end");
        } else {
            c.kind = 'x'; c.nodes = nodes; c.model = m; c.params = kv;   // subcircuit instance, resolved by the flattener
        }
        cur->cards.push_back(c);
    }
}

// ---- flattening ------------------------------------------------------------------------------------------------------
struct Flat {
    // unknowns
    std::vector<std::string> node_names, branch_names;
    std::map<std::string, int> node_index;
    std::map<std::string, std::string> aliases;
    // devices / waves (storage the cb_* structs point into)
    std::vector<cb_device> devices;
    std::vector<std::string> device_names;
    std::vector<cb_wave> waves;
    std::vector<std::unique_ptr<std::vector<double>>> wave_t;
    std::vector<std::unique_ptr<std::vector<cb_pref>>> wave_y;
    // parameter columns
    std::vector<std::string> param_names;
    std::vector<std::vector<double>> columns;
    std::vector<int32_t> outputs;
    std::map<std::string, double> options;   // uniform option values (temp, ...)
    int64_t B = 1;

    int node(const std::string& n0) {
        const std::string n = lower(n0);
        if (n == "0" || n == "gnd" || n == "gnd!") return -1;
        auto it = node_index.find(n);
        if (it != node_index.end()) return it->second;
        const int k = (int)node_names.size();
        node_index[n] = k;
        node_names.push_back(n);
        return k;
    }
    cb_pref value(const std::string& tag, const Val& v) {
        cb_pref p{0.0, -1, 0};
        bool same = true;
        for (size_t i = 1; i < v.v.size(); i++) if (!(v.v[i] == v.v[0])) { same = false; break; }
        if (v.uniform() || same) { p.value = v.v[0]; return p; }
        if ((int64_t)v.v.size() != B) throw Error("sweep column of the wrong length for " + tag);
        for (size_t k = 0; k < param_names.size(); k++)
            if (param_names[k] == tag) { columns[k] = v.v; p.col = (int32_t)k; return p; }
        param_names.push_back(tag);
        columns.push_back(v.v);
        p.col = (int32_t)param_names.size() - 1;
        return p;
    }
    int unknown(const std::string& name0) const {
        std::string key;
        for (char c : lower(name0)) if (c != ' ') key += c;
        const size_t dot = key.rfind('.');
        const std::string last = dot == std::string::npos ? key : key.substr(dot + 1);
        if (last.compare(0, 5, "node_") == 0) key = (dot == std::string::npos ? std::string() : key.substr(0, dot + 1)) + last.substr(5);
        std::set<std::string> seen;
        while (aliases.count(key) && !seen.count(key)) { seen.insert(key); key = aliases.at(key); }
        auto it = node_index.find(key);
        if (it != node_index.end()) return it->second;
        for (size_t k = 0; k < branch_names.size(); k++) if (branch_names[k] == key) return (int)(node_names.size() + k);
        return -1;
    }
};

class Flattener {
    Netlist& nl;
    Flat& fc;
    std::map<std::string, Val> sweep;
    std::set<std::string> used;

    std::map<std::string, Val> overrides(const std::string& prefix, const std::vector<std::string>& names) {
        std::map<std::string, Val> out;
        for (const std::string& n : names) {
            const std::string key = lower(prefix + n);
            auto it = sweep.find(key);
            if (it != sweep.end()) { used.insert(key); out[lower(n)] = it->second; }
        }
        return out;
    }
    // `nothing` in a SerialSweep (NaN here) keeps the default of that point
    static void fill_defaults(Scope& sc, const KV& exprs, const std::map<std::string, Val>& over) {
        for (auto& kv : over) {
            bool any_nan = false;
            for (double x : kv.second.v) any_nan |= std::isnan(x);
            if (!any_nan || !kv_get(exprs, kv.first)) continue;
            sc.values.erase(kv.first);
            Val dflt;
            sc.lookup(kv.first, dflt);
            Val merged = kv.second;
            for (size_t i = 0; i < merged.v.size(); i++) if (std::isnan(merged.v[i])) merged.v[i] = dflt.at(i);
            sc.values[kv.first] = merged;
        }
    }
    static std::map<std::string, std::string> to_map(const KV& kv) {
        std::map<std::string, std::string> m;
        for (auto& p : kv) m[p.first] = p.second;
        return m;
    }
    const Subckt* find_subckt(const Subckt& sub, const std::string& name) const {
        auto it = sub.subckts.find(name);
        if (it != sub.subckts.end()) return it->second.get();
        it = nl.top.subckts.find(name);
        return it == nl.top.subckts.end() ? nullptr : it->second.get();
    }

    int add_wave(const std::string& name, const Source& src, Scope& scope, const std::map<std::string, Val>& over) {
        cb_wave w;
        std::memset(&w, 0, sizeof w);
        for (auto& p : w.v) p = cb_pref{0.0, -1, 0};
        w.dc = cb_pref{0.0, -1, 0};
        bool has_dc = false;
        Val dc;
        if (over.count("dc")) { dc = over.at("dc"); has_dc = true; }
        else if (src.has_dc) { dc = scope.eval(src.dc); has_dc = true; }
        cb_pref dcv{0.0, -1, 0};
        if (has_dc) dcv = fc.value(name + ".dc", dc);
        if (src.tran_kind.empty()) {
            w.kind = CB_W_DC; w.has_dc = 1; w.dc = dcv;   // Wave(W_DC, dc = 0 if none)
        } else {
            std::vector<Val> vals;
            for (const std::string& a : src.tran_args) vals.push_back(scope.eval(a));
            w.has_dc = has_dc ? 1 : 0;
            w.dc = dcv;
            if (src.tran_kind == "pwl") {
                if (vals.size() % 2) throw Error("PWL must have an equal number of x and y values");
                auto ts = std::make_unique<std::vector<double>>();
                auto ys = std::make_unique<std::vector<cb_pref>>();
                for (size_t k = 0; k + 1 < vals.size(); k += 2) {
                    if (!vals[k].uniform()) throw Error(name + ": PWL times cannot be swept (they are shared breakpoints)");
                    ts->push_back(vals[k].v[0]);
                    ys->push_back(fc.value(name + ".pwl" + std::to_string(k / 2), vals[k + 1]));
                }
                w.kind = CB_W_PWL; w.npts = (int32_t)ts->size();
                w.t = ts->data(); w.y = ys->data();
                fc.wave_t.push_back(std::move(ts)); fc.wave_y.push_back(std::move(ys));
            } else {
                const bool pulse = src.tran_kind == "pulse";
                if (vals.size() > 7) throw Error(name + ": too many " + src.tran_kind + " arguments");
                for (size_t k = 0; k < vals.size(); k++) w.v[k] = fc.value(name + "." + (pulse ? "pulse" : "sin") + std::to_string(k), vals[k]);
                w.kind = pulse ? CB_W_PULSE : CB_W_SIN;
                const double inf = std::numeric_limits<double>::infinity();
                if (pulse) {   // PULSE v1 v2 td tr tf [pw [per]]; a period <= 0 means one pulse, as in SPICE
                    if (vals.size() < 5) throw Error(name + ": PULSE needs v1 v2 td tr tf [pw [per]]");
                    for (size_t k = vals.size(); k < 7; k++) w.v[k] = cb_pref{inf, -1, 0};
                    for (int k = 2; k < 7; k++)
                        if (w.v[k].col >= 0) throw Error(name + ": PULSE timing parameters cannot be swept (shared breakpoints)");
                    if (!(w.v[6].value > 0.0)) w.v[6].value = inf;
                    for (int k = 2; k < 6; k++)
                        if (!(w.v[k].value >= 0.0 && w.v[k].value < inf)) throw Error(name + ": PULSE td / tr / tf / pw must be finite and >= 0");
                } else {       // SIN vo va freq td theta phase ncycles
                    const double dflt[7] = {0.0, 0.0, 1.0, 0.0, 0.0, 0.0, inf};
                    for (size_t k = vals.size(); k < 7; k++) w.v[k] = cb_pref{dflt[k], -1, 0};
                }
            }
        }
        if (src.has_ac) {
            const Val ac = scope.eval(src.ac);
            if (!ac.uniform()) throw Error("the AC magnitude of a source cannot be swept");
            w.ac_mag = std::fabs(ac.v[0]);
        }
        fc.waves.push_back(w);
        return (int)fc.waves.size() - 1;
    }

    void add_dev(int kind, const std::string& name, const std::vector<std::string>& nodes, cb_pref value, int wave, double mult, bool branch) {
        cb_device d;
        std::memset(&d, 0, sizeof d);
        d.kind = kind;
        for (int k = 0; k < 4; k++) d.n[k] = k < (int)nodes.size() ? fc.node(nodes[k]) : -1;
        d.branch = branch ? -2 : -1;
        d.wave = wave;
        d.value = value;
        d.mult = mult;
        fc.devices.push_back(d);
        fc.device_names.push_back(name);
    }

    void instantiate(const Subckt& sub, Scope& scope, const std::string& prefix, const std::map<std::string, std::string>& portmap, double mult_ctx) {
        for (auto& pm : portmap) fc.aliases[lower(prefix + pm.first)] = lower(pm.second);
        auto net = [&](const std::string& n) -> std::string {
            if (n == "0" || n == "gnd" || n == "gnd!") return "0";
            auto it = portmap.find(n);
            if (it != portmap.end()) return it->second;
            return prefix + n;
        };
        for (const Card& card : sub.cards) {
            const std::string name = prefix + card.name;
            std::vector<std::string> onames;
            for (auto& p : card.params) onames.push_back(p.first);
            for (const char* s : {"r", "c", "l", "dc", "gain", "w", "nfin", "m"}) onames.push_back(s);
            const std::map<std::string, Val> over = overrides(name + ".", onames);
            auto par = [&](const std::string& key, Val& out) -> bool {
                auto it = over.find(key);
                if (it != over.end()) { out = it->second; return true; }
                if (const std::string* e = kv_get(card.params, key)) { out = scope.eval(*e); return true; }
                return false;
            };
            Val mv(1.0);
            par("m", mv);
            if (!mv.uniform()) throw Error("multiplicity m cannot be swept");
            const double own_m = mv.v[0], mult = own_m * mult_ctx;
            std::vector<std::string> nodes;
            for (const std::string& n : card.nodes) nodes.push_back(net(n));
            const char k = card.kind;
            if (k == 'r' || k == 'c' || k == 'l') {
                const std::string key(1, k);
                Val v;
                bool have = false;
                if (over.count(key)) { v = over.at(key); have = true; }
                else if (card.has_value) { v = scope.eval(card.value); have = true; }
                else have = par(key, v);
                if (!have && !card.model.empty() && !nl.models.count(card.model)) {
                    v = scope.eval(card.model);   // `R2 vcc 0 res`: a bare identifier that is a parameter, not a model
                    have = true;
                }
                if (!have && k == 'r' && !card.model.empty()) {
                    // model card / geometry: r = rsh (l - short) / (w - narrow)  (src/simpledevices.jl:62-70)
                    const std::map<std::string, double>& mp = nl.models.at(card.model);
                    auto g = [&](const char* key2, double d) -> Val {
                        Val x;
                        if (par(key2, x)) return x;
                        auto it = mp.find(key2);
                        return Val(it == mp.end() ? d : it->second);
                    };
                    if (mp.count("r")) v = Val(mp.at("r"));
                    else {
                        const Val num = map2(g("rsh", 50.0), map2(g("l", 1e-6), g("short", 0.0), [](double a, double b) { return a - b; }),
                                             [](double a, double b) { return a * b; });
                        v = map2(num, map2(g("w", 1e-6), g("narrow", 0.0), [](double a, double b) { return a - b; }),
                                 [](double a, double b) { return a / b; });
                    }
                    have = true;
                }
                if (!have) {
                    if (k == 'c') v = Val(1.0);
                    else throw Error(name + ": no value");
                }
                add_dev(k == 'r' ? CB_DEV_R : k == 'c' ? CB_DEV_C : CB_DEV_L, name, nodes, fc.value(name + "." + key, v), -1, mult, k == 'l');
            } else if (k == 'v' || k == 'i') {
                const int w = add_wave(name, card.src, scope, over);
                add_dev(k == 'v' ? CB_DEV_VSRC : CB_DEV_ISRC, name, nodes, cb_pref{0.0, -1, 0}, w, mult, k == 'v');
            } else if (k == 'e' || k == 'g') {
                Val g(1.0);
                if (over.count("gain")) g = over.at("gain");
                else if (card.has_value) g = scope.eval(card.value);
                add_dev(k == 'e' ? CB_DEV_VCVS : CB_DEV_VCCS, name, nodes, fc.value(name + ".gain", g), -1, mult, k == 'e');
            } else if (k == 'x') {
                const Subckt* child = find_subckt(sub, card.model);
                if (!child) throw Error("unknown subcircuit '" + card.model + "' for " + name + " (Verilog-A modules are handled by the Python front end)");
                // instance parameters may refer to one another and to the caller's scope; a name inside its own value
                // means the caller's (`foo=foo+1`)
                Scope inst_scope;
                for (auto& p : card.params) if (p.first != "m") inst_scope.exprs[p.first] = p.second;
                inst_scope.parent = &scope;
                std::map<std::string, Val> given;
                for (auto& p : card.params) if (p.first != "m") { Val v; inst_scope.lookup(p.first, v); given[p.first] = v; }
                std::vector<std::string> cnames;
                for (auto& p : child->params) cnames.push_back(p.first);
                for (auto& p : child->local_params) cnames.push_back(p.first);
                for (auto& o : overrides(name + ".", cnames)) given[o.first] = o.second;
                KV exprs = child->params;
                for (auto& p : child->local_params) {
                    bool found = false;
                    for (auto& q : exprs) if (q.first == p.first) { q.second = p.second; found = true; }
                    if (!found) exprs.push_back(p);
                }
                Scope cs;
                cs.exprs = to_map(exprs);
                cs.parent = &scope;
                cs.values = given;
                fill_defaults(cs, exprs, given);
                if (nodes.size() != child->ports.size())
                    throw Error(name + ": " + std::to_string(nodes.size()) + " nodes for subcircuit " + child->name + " with " +
                                std::to_string(child->ports.size()) + " ports");
                double sub_m = 1.0;
                if (kv_get(card.params, "m") || over.count("m")) sub_m = own_m;
                else if (kv_get(exprs, "m")) { Val mm; cs.lookup("m", mm); if (!mm.uniform()) throw Error("multiplicity m cannot be swept"); sub_m = mm.v[0]; }
                std::map<std::string, std::string> pm;
                for (size_t q = 0; q < nodes.size(); q++) pm[child->ports[q]] = nodes[q];
                instantiate(*child, cs, name + ".", pm, mult_ctx * sub_m);
            } else throw Error("unsupported device " + name);
        }
    }

public:
    Flattener(Netlist& n, Flat& f) : nl(n), fc(f) {}

    void run(const std::vector<std::string>& names, const double* values, int64_t B, const std::vector<std::string>& outputs) {
        fc.B = B;
        for (size_t k = 0; k < names.size(); k++) {
            Val v;
            v.v.assign(values + k * B, values + (k + 1) * B);
            sweep[lower(names[k])] = v;
        }
        std::vector<std::string> top_names;
        for (auto& p : nl.top.params) top_names.push_back(p.first);
        const std::map<std::string, Val> top_over = overrides("", top_names);
        Scope top;
        top.exprs = to_map(nl.top.params);
        top.values = top_over;
        fill_defaults(top, nl.top.params, top_over);
        // `temper` = the simulation temperature in Celsius: a swept `temp`, else `.option temp=` / `.temp`, else 27
        if (sweep.count("temp")) { top.values["temper"] = sweep["temp"]; }
        else if (const std::string* t = kv_get(nl.options, "temp")) top.values["temper"] = Val(parse_number(*t));
        else top.values["temper"] = Val(27.0);
        instantiate(nl.top, top, "", {}, 1.0);
        // branch currents after the node voltages
        const int nn = (int)fc.node_names.size();
        for (size_t d = 0; d < fc.devices.size(); d++)
            if (fc.devices[d].branch == -2) {
                fc.devices[d].branch = nn + (int)fc.branch_names.size();
                fc.branch_names.push_back(fc.device_names[d] + ".i");
            }
        for (auto& kv : nl.options) {
            try {
                const Val v = top.eval(kv.second);
                if (v.uniform()) fc.options[kv.first] = v.v[0];
            } catch (const Error&) {}
        }
        if (sweep.count("temp")) {
            used.insert("temp");
            const Val& t = sweep["temp"];
            bool same = true;
            for (double x : t.v) same &= x == t.v[0];
            if (same) fc.options["temp"] = t.v[0];
            else { fc.value("temp", t); fc.options["temp"] = std::numeric_limits<double>::quiet_NaN(); }   // swept: the column "temp"
        }
        std::string unused;
        for (auto& kv : sweep) if (!used.count(kv.first)) unused += (unused.empty() ? "" : ", ") + kv.first;
        if (!unused.empty()) throw Error("sweep variable(s) " + unused + " do not name any parameter of the circuit");
        for (const std::string& o : outputs) {
            const int u = fc.unknown(o);
            if (u < 0) throw Error("no unknown named '" + o + "'");
            fc.outputs.push_back(u);
        }
    }
};

}   // namespace sf

// The handle behind cb_netlist_*
struct cb_netlist {
    sf::Netlist nl;
    sf::Flat fc;
    std::vector<double> params;   // [P][B]
    std::vector<std::string> unknown_names;
    cb_flat_circuit flat{};
};
