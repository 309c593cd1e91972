"""Verilog-A preprocessor: `define (object and function-like, multi-line), `undef, `ifdef /
`ifndef / `else / `elsif / `endif, `include, macro expansion, comment removal.

Behavioural counterpart of the reference's VerilogAParser.jl/src/parse/preproc.jl; written
from the Verilog-AMS LRM, not translated.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
STD_INCLUDE = os.path.join(_HERE, "include")

_IDENT = re.compile(r"[A-Za-z_][A-Za-z0-9_$]*")


class PreprocError(Exception):
    pass


class Macro:
    __slots__ = ("name", "args", "body")

    def __init__(self, name: str, args: Optional[List[str]], body: str):
        self.name, self.args, self.body = name, args, body


def strip_comments(text: str) -> str:
    """Remove // and /* */ comments, keep strings and newlines (for line-continuations)."""
    out = []
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c == '"':
            j = i + 1
            while j < n and text[j] != '"':
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1])
            i = j + 1
        elif c == "/" and i + 1 < n and text[i + 1] == "/":
            j = text.find("\n", i)
            j = n if j < 0 else j
            # a line comment ending in a backslash must not swallow the continuation
            if j > i and text[j - 1] == "\\":
                out.append(" \\")
            i = j
        elif c == "/" and i + 1 < n and text[i + 1] == "*":
            j = text.find("*/", i + 2)
            if j < 0:
                raise PreprocError("unterminated block comment")
            out.append(" " + "\n" * text.count("\n", i, j))
            i = j + 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


class Preprocessor:
    def __init__(self, include_paths: Sequence[str] = (), defines: Optional[Dict[str, str]] = None,
                 suppress_defines: Sequence[str] = ()):
        self.include_paths = list(include_paths) + [STD_INCLUDE]
        self.macros: Dict[str, Macro] = {}
        self.suppress = set(suppress_defines)
        self.included: List[str] = []
        for k, v in (defines or {}).items():
            self.macros[k] = Macro(k, None, v)

    # ------------------------------------------------------------------
    def process_file(self, path: str) -> str:
        with open(path, "r", errors="replace") as f:
            text = f.read()
        self.included.append(os.path.abspath(path))
        return self.process_text(text, os.path.dirname(os.path.abspath(path)))

    def _find_include(self, name: str, cwd: str) -> str:
        for d in [cwd] + self.include_paths:
            p = os.path.join(d, name)
            if os.path.exists(p):
                return p
        raise PreprocError(f"cannot find include file {name!r}")

    def process_text(self, text: str, cwd: str = ".") -> str:
        text = strip_comments(text)
        # join line continuations (only meaningful inside `define, harmless elsewhere)
        lines = text.split("\n")
        out: List[str] = []
        pending: List[str] = []  # active lines awaiting expansion with the current macro table

        def flush():
            if pending:
                out.append(self.expand("\n".join(pending)))
                pending.clear()

        # condition stack entries: [active_now, any_branch_taken, parent_active]
        stack: List[List[bool]] = []
        i = 0
        while i < len(lines):
            line = lines[i]
            i += 1
            s = line.lstrip()
            active = all(e[0] for e in stack)
            if s.startswith("`"):
                m = _IDENT.match(s, 1)
                word = m.group(0) if m else ""
                rest = s[m.end():] if m else ""
                if word in ("ifdef", "ifndef"):
                    name = rest.split()[0]
                    cond = (name in self.macros) == (word == "ifdef")
                    stack.append([active and cond, cond, active])
                    continue
                if word == "elsif":
                    name = rest.split()[0]
                    e = stack[-1]
                    cond = (name in self.macros) and not e[1]
                    e[0] = e[2] and cond
                    e[1] = e[1] or cond
                    continue
                if word == "else":
                    e = stack[-1]
                    e[0] = e[2] and not e[1]
                    e[1] = True
                    continue
                if word == "endif":
                    stack.pop()
                    tail = rest.strip()
                    if tail:
                        lines.insert(i, tail)
                    continue
                if not active:
                    continue
                if word == "define":
                    full = rest
                    while full.rstrip().endswith("\\") and i < len(lines):
                        full = full.rstrip()[:-1] + "\n" + lines[i]
                        i += 1
                    flush()
                    self._define(full)
                    continue
                if word == "undef":
                    flush()
                    self.macros.pop(rest.split()[0], None)
                    continue
                if word == "include":
                    mm = re.search(r'"([^"]+)"', rest)
                    if not mm:
                        raise PreprocError(f"bad `include: {line!r}")
                    path = self._find_include(mm.group(1), cwd)
                    flush()
                    out.append(self.process_file(path))
                    continue
                if word in ("timescale", "default_discipline", "default_transition", "resetall",
                            "begin_keywords", "end_keywords", "line", "pragma"):
                    continue
            if not active:
                continue
            pending.append(line)
        flush()
        if stack:
            raise PreprocError("unterminated `ifdef")
        return "\n".join(out)

    def _define(self, text: str):
        text = text.lstrip()
        m = _IDENT.match(text)
        if not m:
            raise PreprocError(f"bad `define: {text!r}")
        name = m.group(0)
        pos = m.end()
        args = None
        if pos < len(text) and text[pos] == "(":  # function-like only if '(' follows immediately
            close = text.index(")", pos)
            args = [a.strip() for a in text[pos + 1:close].split(",") if a.strip()]
            pos = close + 1
        body = text[pos:].strip()
        if name in self.suppress:
            return
        self.macros[name] = Macro(name, args, body)

    # ------------------------------------------------------------------
    def expand(self, text: str, depth: int = 0) -> str:
        if "`" not in text:
            return text
        if depth > 64:
            raise PreprocError("macro recursion too deep")
        out = []
        i, n = 0, len(text)
        while i < n:
            c = text[i]
            if c == '"':
                j = i + 1
                while j < n and text[j] != '"':
                    j += 2 if text[j] == "\\" else 1
                out.append(text[i:j + 1])
                i = j + 1
                continue
            if c != "`":
                out.append(c)
                i += 1
                continue
            m = _IDENT.match(text, i + 1)
            if not m:
                out.append(c)
                i += 1
                continue
            name = m.group(0)
            mac = self.macros.get(name)
            if mac is None:
                raise PreprocError(f"undefined macro `{name}")
            i = m.end()
            if mac.args is None:
                out.append(self.expand(mac.body, depth + 1))
                continue
            # function-like: parse balanced argument list
            j = i
            while j < n and text[j].isspace():
                j += 1
            if j >= n or text[j] != "(":
                raise PreprocError(f"macro `{name} needs arguments")
            args, j = self._parse_args(text, j)
            if len(args) != len(mac.args):
                raise PreprocError(f"macro `{name}: expected {len(mac.args)} args, got {len(args)}")
            args = [self.expand(a, depth + 1) for a in args]
            body = self._substitute(mac.body, dict(zip(mac.args, args)))
            out.append(self.expand(body, depth + 1))
            i = j
        return "".join(out)

    @staticmethod
    def _parse_args(text: str, start: int) -> Tuple[List[str], int]:
        assert text[start] == "("
        depth, j, cur, args = 0, start, [], []
        n = len(text)
        while j < n:
            c = text[j]
            if c == '"':
                k = j + 1
                while k < n and text[k] != '"':
                    k += 2 if text[k] == "\\" else 1
                cur.append(text[j:k + 1])
                j = k + 1
                continue
            if c in "([{":
                depth += 1
                if depth > 1:
                    cur.append(c)
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    args.append("".join(cur).strip())
                    return args, j + 1
                cur.append(c)
            elif c == "," and depth == 1:
                args.append("".join(cur).strip())
                cur = []
            else:
                cur.append(c)
            j += 1
        raise PreprocError("unterminated macro argument list (arguments must be on one logical line)")

    @staticmethod
    def _substitute(body: str, mapping: Dict[str, str]) -> str:
        def repl(m):
            return mapping.get(m.group(0), m.group(0))

        # do not substitute inside strings
        parts = re.split(r'("(?:[^"\\]|\\.)*")', body)
        for k in range(0, len(parts), 2):
            parts[k] = _IDENT.sub(repl, parts[k])
        return "".join(parts)


def preprocess_file(path: str, include_paths: Sequence[str] = (), defines=None, suppress_defines=()) -> str:
    pp = Preprocessor(include_paths, defines, suppress_defines)
    return pp.process_file(path)


def preprocess_text(text: str, include_paths: Sequence[str] = (), defines=None, suppress_defines=()) -> str:
    pp = Preprocessor(include_paths, defines, suppress_defines)
    return pp.process_text(text)
