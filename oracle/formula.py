"""Formula evaluator -- ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Restates the arithmetic of the `fasteval` crate, version 0.2.4
(Cargo.lock:636-637 of the reference; the crate source is NOT vendored under
/root/reference, so this follows its published grammar/semantics), as it is
used by the reference at /root/reference/src/gui/uniform.rs:602-635 (parse +
compile) and :1009-1140 (evaluation with the custom-function callback).

Semantics reproduced:
  * precedence, tightest first:  ^  %  /  *  -  +  (== != < <= >= >)  and  or
    -- every arithmetic operator sits on its own level (fasteval README);
  * `^` is right-associative, `%` left-associative (Rust f64 %, sign of dividend);
  * `a / b / c`  is evaluated as  a * (1/b) * (1/c)   (compile: inv-wrap + mul);
  * `a - b - c`  is evaluated as  a + (-b) + (-c)     (compile: neg-wrap + add);
  * in a sum / product literal constants are folded together and applied last;
  * unary +, -, ! bind to the following value only;
  * built-in functions: int ceil floor abs sign log round min max e pi sin cos
    tan asin acos atan sinh cosh tanh asinh acosh atanh; any other name (with
    or without arguments) goes to the caller's namespace callback -- exactly
    like the reference's `cb` closure (uniform.rs:1014-1124).

parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md);
the pins are this repo's committed goldens and its second, independent implementations.
"""
from __future__ import annotations

import math

_EPS8 = 8.0 * 2.220446049250313e-16


class FormulaError(ValueError):
    pass


# ----------------------------------------------------------------- tokenizer
def _tokenize(text):
    toks = []
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c.isspace():
            i += 1
        elif c.isdigit() or (c == "." and i + 1 < n and text[i + 1].isdigit()):
            j = i
            while j < n and (text[j].isdigit() or text[j] == "."):
                j += 1
            if j < n and text[j] in "eE":
                k = j + 1
                if k < n and text[k] in "+-":
                    k += 1
                if k < n and text[k].isdigit():
                    j = k
                    while j < n and text[j].isdigit():
                        j += 1
            try:
                toks.append(("num", float(text[i:j])))
            except ValueError:
                raise FormulaError(f"malformed number {text[i:j]!r} in formula {text!r}") from None
            i = j
        elif c.isalpha() or c == "_":
            j = i
            while j < n and (text[j].isalnum() or text[j] in "_."):
                j += 1
            toks.append(("id", text[i:j]))
            i = j
        else:
            two = text[i:i + 2]
            if two in ("==", "!=", "<=", ">=", "&&", "||"):
                toks.append(("op", two))
                i += 2
            elif c in "+-*/%^<>!(),[]":
                toks.append(("op", c))
                i += 1
            else:
                raise FormulaError(f"unexpected character {c!r} in formula {text!r}")
    toks.append(("end", None))
    return toks


# Binary operator priority (higher binds tighter), as fasteval orders BinaryOp.
_PRIO = {"or": 1, "and": 2, "cmp": 3, "+": 4, "-": 5, "*": 6, "/": 7, "%": 8, "^": 9}
_CMP = ("==", "!=", "<", "<=", ">=", ">")


def _binop_of(tok):
    kind, val = tok
    if kind == "op":
        if val in ("+", "-", "*", "/", "%", "^") or val in _CMP:
            return val
        if val == "&&":
            return "and"
        if val == "||":
            return "or"
    if kind == "id" and val in ("and", "or"):
        return val
    return None


_BUILTIN = {
    "int", "ceil", "floor", "abs", "sign", "log", "round", "min", "max", "e", "pi",
    "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
    "asinh", "acosh", "atanh",
}


class _Parser:
    def __init__(self, text):
        self.text = text
        self.toks = _tokenize(text)
        self.i = 0

    def peek(self):
        return self.toks[self.i]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def expression(self):
        """-> ('expr', first_value, [(op, value), ...])  (flat, like fasteval's Expression)"""
        first = self.value()
        pairs = []
        while True:
            op = _binop_of(self.peek())
            if op is None:
                break
            self.next()
            pairs.append((op, self.value()))
        return ("expr", first, pairs)

    def value(self):
        kind, val = self.peek()
        if kind == "num":
            self.next()
            return ("const", val)
        if kind == "op" and val in "+-!":
            self.next()
            inner = self.value()
            return ({"+": "pos", "-": "neg", "!": "not"}[val], inner)
        if kind == "op" and val in "([":
            self.next()
            e = self.expression()
            k2, v2 = self.next()
            if not (k2 == "op" and v2 in ")]"):
                raise FormulaError(f"expected ) in {self.text!r}")
            return e
        if kind == "id":
            self.next()
            args = None
            k2, v2 = self.peek()
            if k2 == "op" and v2 in "([":
                self.next()
                args = []
                k3, v3 = self.peek()
                if k3 == "op" and v3 in ")]":
                    self.next()
                else:
                    while True:
                        args.append(self.expression())
                        k3, v3 = self.next()
                        if k3 == "op" and v3 in ")]":
                            break
                        if not (k3 == "op" and v3 == ","):
                            raise FormulaError(f"expected , or ) in {self.text!r}")
            return ("call", val, args)
        raise FormulaError(f"unexpected token {val!r} in {self.text!r}")


def parse(text):
    p = _Parser(text)
    e = p.expression()
    if p.peek()[0] != "end":
        raise FormulaError(f"trailing tokens in {text!r}")
    return e


# ---------------------------------------------------------------- evaluation
def _is_const(node):
    return node[0] == "const"


def _split(first, pairs, op):
    """Split a flat expression at every occurrence of `op` (cmp: any comparison)."""
    groups = []
    cur_first, cur_pairs = first, []
    ops_used = []
    for o, v in pairs:
        hit = (o in _CMP) if op == "cmp" else (o == op)
        if hit:
            groups.append((cur_first, cur_pairs))
            ops_used.append(o)
            cur_first, cur_pairs = v, []
        else:
            cur_pairs.append((o, v))
    groups.append((cur_first, cur_pairs))
    return groups, ops_used


def _prio(o):
    return _PRIO["cmp"] if o in _CMP else _PRIO[o]


def _fmod(a, b):
    if b == 0.0 or math.isinf(a) or math.isnan(a) or math.isnan(b):
        return math.nan
    return math.fmod(a, b)


def _pow(a, b):
    """f64::powf (libm pow) without Python's exceptions."""
    try:
        return math.pow(a, b)
    except OverflowError:                       # finite operands, result out of range: +-inf, negative for (-x)^odd
        odd = math.isfinite(b) and b == math.floor(b) and math.fmod(b, 2.0) != 0.0
        return -math.inf if (a < 0 and odd) else math.inf
    except ValueError:
        if a == 0.0 and b < 0:                  # pow(+-0, negative) = +-inf (pole), Python calls it a domain error
            odd = math.isfinite(b) and b == math.floor(b) and math.fmod(b, 2.0) != 0.0
            return math.copysign(math.inf, a) if odd else math.inf
        return math.nan                         # negative base, non-integral exponent


def _div(a, b):
    if b == 0.0:
        if a == 0.0 or math.isnan(a):
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


class Evaluator:
    """ns(name, args: list[float]) -> float | None  (None => evaluation fails)."""

    def __init__(self, ns):
        self.ns = ns

    def eval_text(self, text):
        return self.eval(parse(text))

    def eval(self, node):
        k = node[0]
        if k == "const":
            return node[1]
        if k == "pos":
            return self.eval(node[1])
        if k == "neg":
            return -self.eval(node[1])
        if k == "not":
            return 1.0 if abs(self.eval(node[1])) <= _EPS8 else 0.0
        if k == "call":
            return self.call(node[1], node[2])
        if k == "expr":
            return self.eval_slice(node[1], node[2])
        raise FormulaError(f"bad node {k}")

    def eval_slice(self, first, pairs):
        if not pairs:
            return self.eval(first)
        lowest = min(pairs, key=lambda p: _prio(p[0]))[0]
        if lowest in _CMP:
            groups, ops = _split(first, pairs, "cmp")
            out = self.eval_slice(*groups[0])
            for o, g in zip(ops, groups[1:]):
                r = self.eval_slice(*g)
                if o == "==":
                    t = abs(out - r) <= _EPS8
                elif o == "!=":
                    t = abs(out - r) > _EPS8
                elif o == "<":
                    t = out < r
                elif o == "<=":
                    t = out <= r
                elif o == ">=":
                    t = out >= r
                else:
                    t = out > r
                out = 1.0 if t else 0.0
            return out
        groups, _ = _split(first, pairs, lowest)
        if lowest == "or":
            out = 0.0
            for g in groups:
                out = self.eval_slice(*g)
                if abs(out) > _EPS8:
                    return out
            return out
        if lowest == "and":
            out = 0.0
            for g in groups:
                out = self.eval_slice(*g)
                if abs(out) <= _EPS8:
                    return out
            return out
        if lowest in ("+", "-"):
            const_sum = 0.0
            acc = None
            for idx, g in enumerate(groups):
                is_c = (not g[1]) and _is_const(g[0])
                v = self.eval_slice(*g)
                if lowest == "-" and idx > 0:
                    v = -v
                if is_c:
                    const_sum += v
                else:
                    acc = v if acc is None else acc + v
            if acc is None:
                return const_sum
            if const_sum != 0.0:
                acc = acc + const_sum
            return acc
        if lowest in ("*", "/"):
            const_prod = 1.0
            acc = None
            for idx, g in enumerate(groups):
                is_c = (not g[1]) and _is_const(g[0])
                v = self.eval_slice(*g)
                if lowest == "/" and idx > 0:
                    v = _div(1.0, v)
                if is_c:
                    const_prod *= v
                else:
                    acc = v if acc is None else acc * v
            if acc is None:
                return const_prod
            if const_prod != 1.0:
                acc = acc * const_prod
            return acc
        if lowest == "%":
            out = self.eval_slice(*groups[0])
            for g in groups[1:]:
                out = _fmod(out, self.eval_slice(*g))
            return out
        if lowest == "^":
            out = self.eval_slice(*groups[-1])
            for g in reversed(groups[:-1]):
                out = _pow(self.eval_slice(*g), out)
            return out
        raise FormulaError(f"bad operator {lowest}")

    def call(self, name, arg_nodes):
        if arg_nodes is not None and name in _BUILTIN:
            a = [self.eval(x) for x in arg_nodes]
            return _builtin(name, a)
        args = [] if arg_nodes is None else [self.eval(x) for x in arg_nodes]
        r = self.ns(name, args)
        if r is None:
            raise FormulaError(f"undefined name {name!r}")
        return float(r)


def _builtin(name, a):
    def need(n):
        if len(a) < n:
            raise FormulaError(f"{name} needs {n} argument(s)")

    if name == "pi":
        return math.pi
    if name == "e":
        return math.e
    if name in ("min", "max"):
        need(1)
        out = a[0]
        for v in a[1:]:
            # f64::min / max semantics: a NaN operand is ignored; among equal values (-0.0 vs 0.0) the earlier argument stays
            if math.isnan(out):
                out = v
            elif not math.isnan(v) and (v < out if name == "min" else v > out):
                out = v
        return out
    if name == "log":
        need(1)
        base, x = (10.0, a[0]) if len(a) == 1 else (a[0], a[1])

        def ieee_log(fn, v):                       # Rust f64::ln / log2 / log10: ln(0) = -inf, ln(negative) = NaN
            if math.isnan(v) or v < 0.0:
                return math.nan
            if v == 0.0:
                return -math.inf
            return fn(v)
        if base == 2.0:
            return ieee_log(math.log2, x)
        if base == 10.0:
            return ieee_log(math.log10, x)
        return _div(ieee_log(math.log, x), ieee_log(math.log, base))
    if name == "round":
        need(1)
        modulus, x = (1.0, a[0]) if len(a) == 1 else (a[0], a[1])
        q = _div(x, modulus)
        if math.isnan(q) or math.isinf(q):
            r = q
        else:
            fl = float(math.floor(abs(q)))                             # Rust round: half away from zero, sign of q kept;
            r = math.copysign(fl + 1.0 if abs(q) - fl >= 0.5 else fl, q)  # (not floor(|q| + 0.5): 0.49999999999999994 must give 0)
        return r * modulus
    need(1)
    x = a[0]
    # Rust's f64 methods are total (IEEE): spell out what Python's math module turns into exceptions
    if math.isnan(x):
        return math.nan
    if math.isinf(x):
        if name in ("int", "ceil", "floor", "sinh", "asinh"):
            return x
        if name in ("abs", "cosh"):
            return math.inf
        if name in ("sign", "tanh"):
            return math.copysign(1.0, x)
        if name == "atan":
            return math.copysign(math.pi / 2.0, x)
        if name == "acosh":
            return math.inf if x > 0 else math.nan
        return math.nan                                               # sin cos tan asin acos atanh of +-inf
    try:
        if name in ("int", "ceil", "floor"):
            r = float({"int": math.trunc, "ceil": math.ceil, "floor": math.floor}[name](x))
            return math.copysign(0.0, x) if r == 0.0 else r        # IEEE keeps the sign of a zero result (ceil(-0.3) = -0.0)
        if name == "abs":
            return abs(x)
        if name == "sign":
            return math.copysign(1.0, x)  # Rust f64::signum
        if name == "atanh" and abs(x) == 1.0:
            return math.copysign(math.inf, x)
        return getattr(math, name)(x)
    except OverflowError:                                               # sinh / cosh beyond the f64 range
        return math.inf if name == "cosh" or x > 0 else -math.inf
    except ValueError:
        return math.nan
