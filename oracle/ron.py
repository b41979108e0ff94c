"""Minimal RON (Rusty Object Notation) reader -- ORACLE / TEST INFRASTRUCTURE ONLY.

Part of the CPU oracle (see oracle/README.md).  Nothing under portal_b200/ may
import this module; the product has its own RON reader in C++
(portal_b200/csrc/host/ron.cpp).

Reads the subset of RON 0.10 that the reference's scene files use
(/root/reference/src/gui/scene_serialized.rs:611-646 is the schema; the files
are written by ron::ser::PrettyConfig with escape_strings(false),
scene_serialized.rs:22-24).

Mapping to Python values:
    (a: 1, b: 2)        -> dict                          (struct)
    (1, 2)              -> list                          (tuple / newtype)
    Name(a: 1)          -> Tagged("Name", dict)          (struct variant)
    Name(1, 2)          -> Tagged("Name", list)          (tuple variant)
    Name                -> Tagged("Name", None)          (unit variant)
    Some(x) / None      -> x / None
    [..]                -> list,  {k: v} -> dict
    "..", r#".."#       -> str; numbers -> int | float; true/false -> bool

parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md);
the pins are this repo's committed goldens and its second, independent implementations.
"""
from __future__ import annotations


class Tagged:
    __slots__ = ("tag", "value")

    def __init__(self, tag, value):
        self.tag = tag
        self.value = value

    def __repr__(self):
        return f"{self.tag}({self.value!r})"

    def __eq__(self, other):
        return isinstance(other, Tagged) and self.tag == other.tag and self.value == other.value


class RonError(ValueError):
    pass


_IDENT_START = set("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ_")
_IDENT_CONT = _IDENT_START | set("0123456789")
_NUM_START = set("0123456789+-.")


class _Parser:
    def __init__(self, text: str):
        self.s = text
        self.i = 0
        self.n = len(text)

    def err(self, msg):
        line = self.s.count("\n", 0, self.i) + 1
        raise RonError(f"RON parse error at line {line}: {msg}")

    def ws(self):
        s, n = self.s, self.n
        while self.i < n:
            c = s[self.i]
            if c in " \t\r\n":
                self.i += 1
            elif c == "/" and self.i + 1 < n and s[self.i + 1] == "/":
                j = s.find("\n", self.i)
                self.i = n if j < 0 else j + 1
            elif c == "/" and self.i + 1 < n and s[self.i + 1] == "*":
                j = s.find("*/", self.i + 2)
                if j < 0:
                    self.err("unterminated block comment")
                self.i = j + 2
            else:
                break

    def peek(self):
        self.ws()
        return self.s[self.i] if self.i < self.n else ""

    def expect(self, ch):
        if self.peek() != ch:
            self.err(f"expected {ch!r}, found {self.s[self.i:self.i+10]!r}")
        self.i += 1

    def ident(self):
        j = self.i
        while j < self.n and self.s[j] in _IDENT_CONT:
            j += 1
        out = self.s[self.i:j]
        self.i = j
        return out

    def string(self):
        # self.s[self.i] == '"'
        s = self.s
        j = self.i + 1
        out = []
        while True:
            if j >= self.n:
                self.err("unterminated string")
            c = s[j]
            if c == '"':
                break
            if c == "\\":
                e = s[j + 1]
                if e == "n":
                    out.append("\n")
                elif e == "t":
                    out.append("\t")
                elif e == "r":
                    out.append("\r")
                elif e == "0":
                    out.append("\0")
                elif e in "\\\"'":
                    out.append(e)
                elif e == "u":
                    k = s.index("}", j)
                    out.append(chr(int(s[j + 3:k], 16)))
                    j = k - 1
                else:
                    self.err(f"unknown escape \\{e}")
                j += 2
                continue
            out.append(c)
            j += 1
        self.i = j + 1
        return "".join(out)

    def raw_string(self):
        # r"..." or r#"..."#
        j = self.i + 1
        hashes = 0
        while self.s[j] == "#":
            hashes += 1
            j += 1
        if self.s[j] != '"':
            self.err("bad raw string")
        end = '"' + "#" * hashes
        k = self.s.find(end, j + 1)
        if k < 0:
            self.err("unterminated raw string")
        out = self.s[j + 1:k]
        self.i = k + len(end)
        return out

    def number(self):
        j = self.i
        s = self.s
        if s[j] in "+-":
            j += 1
        while j < self.n and (s[j].isdigit() or s[j] in "._eE" or (s[j] in "+-" and s[j - 1] in "eE")):
            j += 1
        tok = s[self.i:j].replace("_", "")
        self.i = j
        if tok in ("+", "-") and s.startswith("inf", j):
            self.i = j + 3
            return float(tok + "inf")
        try:
            if any(c in tok for c in ".eE"):
                return float(tok)
            return int(tok)
        except ValueError:
            self.err(f"bad number {tok!r}")

    def seq_or_struct(self):
        """Parse the inside of '(' ... ')' -> dict (if 'ident:' fields) or list."""
        self.expect("(")
        # Lookahead: struct field = identifier followed by ':'
        save = self.i
        is_struct = False
        if self.peek() in _IDENT_START:
            self.ident()
            if self.peek() == ":":
                is_struct = True
        self.i = save
        if is_struct:
            out = {}
            while True:
                if self.peek() == ")":
                    self.i += 1
                    return out
                key = self.ident()
                self.expect(":")
                out[key] = self.value()
                if self.peek() == ",":
                    self.i += 1
        out = []
        while True:
            if self.peek() == ")":
                self.i += 1
                return out
            out.append(self.value())
            if self.peek() == ",":
                self.i += 1

    def value(self):
        c = self.peek()
        if c == "":
            self.err("unexpected end of input")
        if c == "(":
            return self.seq_or_struct()
        if c == "[":
            self.i += 1
            out = []
            while True:
                if self.peek() == "]":
                    self.i += 1
                    return out
                out.append(self.value())
                if self.peek() == ",":
                    self.i += 1
        if c == "{":
            self.i += 1
            out = {}
            while True:
                if self.peek() == "}":
                    self.i += 1
                    return out
                k = self.value()
                self.expect(":")
                out[k] = self.value()
                if self.peek() == ",":
                    self.i += 1
        if c == '"':
            return self.string()
        if c == "r" and self.i + 1 < self.n and self.s[self.i + 1] in '"#':
            return self.raw_string()
        if c in _IDENT_START:
            name = self.ident()
            if name == "true":
                return True
            if name == "false":
                return False
            if name == "inf":
                return float("inf")
            if name == "NaN":
                return float("nan")
            if self.peek() == "(":
                inner = self.seq_or_struct()
                if name == "Some":
                    if not isinstance(inner, list) or len(inner) != 1:
                        self.err("Some(..) takes one value")
                    return inner[0]
                return Tagged(name, inner)
            if name == "None":
                return None
            return Tagged(name, None)
        if c in _NUM_START:
            return self.number()
        self.err(f"unexpected character {c!r}")


def loads(text: str):
    p = _Parser(text)
    v = p.value()
    if p.peek() != "":
        p.err("trailing characters")
    return v


def load(path: str):
    with open(path, "r", encoding="utf-8") as f:
        return loads(f.read())
