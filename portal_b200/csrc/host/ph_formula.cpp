#include "ph_formula.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace ph {

struct FormulaNode {
    enum Kind { Const, Pos, Neg, Not, Call, Expr } kind = Const;
    double value = 0.0;                     // Const
    FormulaPtr child;                       // Pos / Neg / Not
    std::string name;                       // Call
    bool has_args = false;                  // Call: `f(...)` vs bare `f`
    std::vector<FormulaPtr> args;           // Call
    FormulaPtr first;                       // Expr: first value
    std::vector<std::pair<int, FormulaPtr>> pairs;  // Expr: (binary op, value)...
};

namespace {

enum Op { OR = 1, AND, EQ, NE, LT, LE, GE, GT, ADD, SUB, MUL, DIV, MOD, EXP };
inline bool is_cmp(int o) { return o >= EQ && o <= GT; }
// priority: higher binds tighter (fasteval's BinaryOp ordering)
inline int prio(int o) {
    switch (o) {
        case OR: return 1;
        case AND: return 2;
        case ADD: return 4;
        case SUB: return 5;
        case MUL: return 6;
        case DIV: return 7;
        case MOD: return 8;
        case EXP: return 9;
        default: return 3;  // comparisons share one level
    }
}

const double EPS8 = 8.0 * 2.220446049250313e-16;

struct Tok {
    enum T { Num, Id, Sym, End } t = End;
    double num = 0;
    std::string s;
};

struct Lexer {
    const std::string& text;
    size_t i = 0;
    explicit Lexer(const std::string& t) : text(t) {}
    Tok next() {
        while (i < text.size() && std::isspace((unsigned char)text[i])) i++;
        Tok k;
        if (i >= text.size()) return k;
        char c = text[i];
        if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < text.size() && std::isdigit((unsigned char)text[i + 1]))) {
            size_t j = i;
            while (j < text.size() && (std::isdigit((unsigned char)text[j]) || text[j] == '.')) j++;
            if (j < text.size() && (text[j] == 'e' || text[j] == 'E')) {
                size_t q = j + 1;
                if (q < text.size() && (text[q] == '+' || text[q] == '-')) q++;
                if (q < text.size() && std::isdigit((unsigned char)text[q])) {
                    j = q;
                    while (j < text.size() && std::isdigit((unsigned char)text[j])) j++;
                }
            }
            k.t = Tok::Num;
            k.num = std::strtod(text.substr(i, j - i).c_str(), nullptr);
            i = j;
            return k;
        }
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t j = i;
            while (j < text.size() && (std::isalnum((unsigned char)text[j]) || text[j] == '_' || text[j] == '.')) j++;
            k.t = Tok::Id;
            k.s = text.substr(i, j - i);
            i = j;
            return k;
        }
        static const char* two[] = {"==", "!=", "<=", ">=", "&&", "||"};
        for (const char* t : two)
            if (text.compare(i, 2, t) == 0) {
                k.t = Tok::Sym;
                k.s = t;
                i += 2;
                return k;
            }
        if (std::string("+-*/%^<>!(),[]").find(c) != std::string::npos) {
            k.t = Tok::Sym;
            k.s = std::string(1, c);
            i++;
            return k;
        }
        throw std::runtime_error(std::string("unexpected character `") + c + "` in formula");
    }
};

struct Parser {
    Lexer lex;
    Tok cur;
    explicit Parser(const std::string& t) : lex(t) { cur = lex.next(); }
    void advance() { cur = lex.next(); }

    int binop() const {
        if (cur.t == Tok::Sym) {
            const std::string& s = cur.s;
            if (s == "+") return ADD;
            if (s == "-") return SUB;
            if (s == "*") return MUL;
            if (s == "/") return DIV;
            if (s == "%") return MOD;
            if (s == "^") return EXP;
            if (s == "==") return EQ;
            if (s == "!=") return NE;
            if (s == "<") return LT;
            if (s == "<=") return LE;
            if (s == ">=") return GE;
            if (s == ">") return GT;
            if (s == "&&") return AND;
            if (s == "||") return OR;
        }
        if (cur.t == Tok::Id && cur.s == "and") return AND;
        if (cur.t == Tok::Id && cur.s == "or") return OR;
        return 0;
    }

    FormulaPtr expression() {
        auto e = std::make_shared<FormulaNode>();
        e->kind = FormulaNode::Expr;
        e->first = value();
        while (int op = binop()) {
            advance();
            e->pairs.emplace_back(op, value());
        }
        return e;
    }

    bool close_bracket() const { return cur.t == Tok::Sym && (cur.s == ")" || cur.s == "]"); }
    bool open_bracket() const { return cur.t == Tok::Sym && (cur.s == "(" || cur.s == "["); }

    FormulaPtr value() {
        auto n = std::make_shared<FormulaNode>();
        if (cur.t == Tok::Num) {
            n->kind = FormulaNode::Const;
            n->value = cur.num;
            advance();
            return n;
        }
        if (cur.t == Tok::Sym && (cur.s == "+" || cur.s == "-" || cur.s == "!")) {
            n->kind = cur.s == "+" ? FormulaNode::Pos : (cur.s == "-" ? FormulaNode::Neg : FormulaNode::Not);
            advance();
            n->child = value();
            return n;
        }
        if (open_bracket()) {
            advance();
            FormulaPtr e = expression();
            if (!close_bracket()) throw std::runtime_error("expected `)` in formula");
            advance();
            return e;
        }
        if (cur.t == Tok::Id) {
            n->kind = FormulaNode::Call;
            n->name = cur.s;
            advance();
            if (open_bracket()) {
                n->has_args = true;
                advance();
                if (close_bracket()) {
                    advance();
                } else {
                    while (true) {
                        n->args.push_back(expression());
                        if (close_bracket()) { advance(); break; }
                        if (!(cur.t == Tok::Sym && cur.s == ",")) throw std::runtime_error("expected `,` or `)` in formula");
                        advance();
                    }
                }
            }
            return n;
        }
        throw std::runtime_error("unexpected token in formula");
    }
};

struct Evaluator {
    const FormulaNamespace& ns;
    explicit Evaluator(const FormulaNamespace& n) : ns(n) {}

    struct Slice {
        const FormulaNode* first;
        const std::pair<int, FormulaPtr>* pairs;
        size_t n;
    };

    double eval(const FormulaNode* node) {
        switch (node->kind) {
            case FormulaNode::Const: return node->value;
            case FormulaNode::Pos: return eval(node->child.get());
            case FormulaNode::Neg: return -eval(node->child.get());
            case FormulaNode::Not: return std::fabs(eval(node->child.get())) <= EPS8 ? 1.0 : 0.0;
            case FormulaNode::Call: return call(node);
            case FormulaNode::Expr: return eval_slice({node->first.get(), node->pairs.data(), node->pairs.size()});
        }
        return 0.0;
    }

    static bool slice_is_const(const Slice& s) { return s.n == 0 && s.first->kind == FormulaNode::Const; }

    // split at every operator of class `op` (op == EQ stands for "any comparison")
    static std::vector<Slice> split(const Slice& s, int op, std::vector<int>* ops_used) {
        std::vector<Slice> out;
        const FormulaNode* first = s.first;
        size_t start = 0;
        for (size_t k = 0; k < s.n; k++) {
            int o = s.pairs[k].first;
            bool hit = op == EQ ? is_cmp(o) : o == op;
            if (hit) {
                out.push_back({first, s.pairs + start, k - start});
                if (ops_used) ops_used->push_back(o);
                first = s.pairs[k].second.get();
                start = k + 1;
            }
        }
        out.push_back({first, s.pairs + start, s.n - start});
        return out;
    }

    double eval_slice(const Slice& s) {
        if (s.n == 0) return eval(s.first);
        int lowest = s.pairs[0].first;
        for (size_t k = 1; k < s.n; k++)
            if (prio(s.pairs[k].first) < prio(lowest)) lowest = s.pairs[k].first;
        if (is_cmp(lowest)) {
            std::vector<int> ops;
            auto groups = split(s, EQ, &ops);
            double out = eval_slice(groups[0]);
            for (size_t k = 1; k < groups.size(); k++) {
                double r = eval_slice(groups[k]);
                bool t = false;
                switch (ops[k - 1]) {
                    case EQ: t = std::fabs(out - r) <= EPS8; break;
                    case NE: t = std::fabs(out - r) > EPS8; break;
                    case LT: t = out < r; break;
                    case LE: t = out <= r; break;
                    case GE: t = out >= r; break;
                    case GT: t = out > r; break;
                }
                out = t ? 1.0 : 0.0;
            }
            return out;
        }
        auto groups = split(s, lowest, nullptr);
        switch (lowest) {
            case OR: {
                double out = 0.0;
                for (auto& g : groups) {
                    out = eval_slice(g);
                    if (std::fabs(out) > EPS8) return out;
                }
                return out;
            }
            case AND: {
                double out = 0.0;
                for (auto& g : groups) {
                    out = eval_slice(g);
                    if (std::fabs(out) <= EPS8) return out;
                }
                return out;
            }
            case ADD:
            case SUB: {
                double const_sum = 0.0, acc = 0.0;
                bool have = false;
                for (size_t k = 0; k < groups.size(); k++) {
                    double v = eval_slice(groups[k]);
                    if (lowest == SUB && k > 0) v = -v;
                    if (slice_is_const(groups[k])) const_sum += v;
                    else {
                        acc = have ? acc + v : v;
                        have = true;
                    }
                }
                if (!have) return const_sum;
                return const_sum != 0.0 ? acc + const_sum : acc;
            }
            case MUL:
            case DIV: {
                double const_prod = 1.0, acc = 0.0;
                bool have = false;
                for (size_t k = 0; k < groups.size(); k++) {
                    double v = eval_slice(groups[k]);
                    if (lowest == DIV && k > 0) v = 1.0 / v;
                    if (slice_is_const(groups[k])) const_prod *= v;
                    else {
                        acc = have ? acc * v : v;
                        have = true;
                    }
                }
                if (!have) return const_prod;
                return const_prod != 1.0 ? acc * const_prod : acc;
            }
            case MOD: {
                double out = eval_slice(groups[0]);
                for (size_t k = 1; k < groups.size(); k++) out = std::fmod(out, eval_slice(groups[k]));
                return out;
            }
            case EXP: {
                double out = eval_slice(groups.back());
                for (size_t k = groups.size() - 1; k-- > 0;) out = std::pow(eval_slice(groups[k]), out);
                return out;
            }
        }
        throw std::runtime_error("bad operator");
    }

    double call(const FormulaNode* n) {
        std::vector<double> a;
        a.reserve(n->args.size());
        for (auto& x : n->args) a.push_back(eval(x.get()));
        double out;
        if (n->has_args && builtin(n->name, a, out)) return out;
        if (!ns(n->name, a, out)) throw std::runtime_error("cannot find variable or function `" + n->name + "`");
        return out;
    }

    static bool builtin(const std::string& f, const std::vector<double>& a, double& out) {
        auto need = [&](size_t k) {
            if (a.size() < k) throw std::runtime_error(f + ": missing argument");
        };
        if (f == "pi") { out = M_PI; return true; }
        if (f == "e") { out = M_E; return true; }
        if (f == "min" || f == "max") {
            need(1);
            out = a[0];
            // f64::min / max semantics: a NaN operand is ignored; among equal values (-0.0 vs 0.0) the earlier argument stays
            for (size_t k = 1; k < a.size(); k++) {
                const double v = a[k];
                if (std::isnan(out)) out = v;
                else if (!std::isnan(v) && (f == "min" ? v < out : v > out)) out = v;
            }
            return true;
        }
        if (f == "log") {
            need(1);
            double base = a.size() == 1 ? 10.0 : a[0], x = a.size() == 1 ? a[0] : a[1];
            out = base == 2.0 ? std::log2(x) : (base == 10.0 ? std::log10(x) : std::log(x) / std::log(base));
            return true;
        }
        if (f == "round") {
            need(1);
            double modulus = a.size() == 1 ? 1.0 : a[0], x = a.size() == 1 ? a[0] : a[1];
            out = std::round(x / modulus) * modulus;  // half away from zero, like Rust's f64::round
            return true;
        }
        static const char* unary[] = {"int", "ceil", "floor", "abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan",
                                      "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"};
        bool known = false;
        for (const char* u : unary)
            if (f == u) known = true;
        if (!known) return false;
        need(1);
        double x = a[0];
        if (f == "int") out = std::trunc(x);
        else if (f == "ceil") out = std::ceil(x);
        else if (f == "floor") out = std::floor(x);
        else if (f == "abs") out = std::fabs(x);
        else if (f == "sign") out = std::isnan(x) ? x : std::copysign(1.0, x);  // f64::signum (NaN stays NaN)
        else if (f == "sin") out = std::sin(x);
        else if (f == "cos") out = std::cos(x);
        else if (f == "tan") out = std::tan(x);
        else if (f == "asin") out = std::asin(x);
        else if (f == "acos") out = std::acos(x);
        else if (f == "atan") out = std::atan(x);
        else if (f == "sinh") out = std::sinh(x);
        else if (f == "cosh") out = std::cosh(x);
        else if (f == "tanh") out = std::tanh(x);
        else if (f == "asinh") out = std::asinh(x);
        else if (f == "acosh") out = std::acosh(x);
        else out = std::atanh(x);
        return true;
    }
};

}  // namespace

FormulaPtr formula_parse(const std::string& text, std::string& err) {
    try {
        Parser p(text);
        FormulaPtr e = p.expression();
        if (p.cur.t != Tok::End) throw std::runtime_error("trailing tokens in formula");
        return e;
    } catch (const std::exception& ex) {
        err = std::string(ex.what()) + " in `" + text + "`";
        return nullptr;
    }
}

bool formula_eval(const FormulaPtr& f, const FormulaNamespace& ns, double& out, std::string& err) {
    if (!f) {
        err = "formula did not parse";
        return false;
    }
    try {
        Evaluator ev(ns);
        out = ev.eval(f.get());
        return true;
    } catch (const std::exception& ex) {
        err = ex.what();
        return false;
    }
}

}  // namespace ph
