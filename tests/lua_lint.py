"""A small Lua 5.1 lexer + structure check (no Lua interpreter exists in this image): tokenises a chunk — long and short
strings, long and line comments, numbers, names, operators — and verifies what a missing `end`, a stray bracket or an
unterminated string would break: block keywords balance (`function` / `if` / `for` / `while` / `do` ... `end`,
`repeat` ... `until`), brackets nest, `then` / `do` follow their openers.  Returns the token list for further checks
(which global names a file calls, which methods a table defines)."""
import re

KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local", "nil", "not",
            "or", "repeat", "return", "then", "true", "until", "while"}
_NAME = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
_NUM = re.compile(r"0[xX][0-9a-fA-F]+(?:[uU]?[lL]{2})?|\d+\.?\d*(?:[eE][+-]?\d+)?(?:[uU]?[lL]{2})?|\.\d+(?:[eE][+-]?\d+)?")
_OPS = ["...", "..", "==", "~=", "<=", ">=", "+", "-", "*", "/", "%", "^", "#", "<", ">", "=", "(", ")", "{", "}", "[", "]",
        ";", ":", ",", "."]


class LuaSyntaxError(Exception):
    pass


def _long_bracket(src, i):
    """if src[i:] opens a long bracket `[==[`, return (level, index after the opener), else None"""
    if src[i] != "[":
        return None
    j = i + 1
    while j < len(src) and src[j] == "=":
        j += 1
    if j < len(src) and src[j] == "[":
        return j - i - 1, j + 1
    return None


def tokenize(src):
    toks, i, line = [], 0, 1
    n = len(src)
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1
            i += 1
        elif c in " \t\r":
            i += 1
        elif src.startswith("--", i):
            lb = _long_bracket(src, i + 2) if i + 2 < n else None
            if lb:
                close = "]" + "=" * lb[0] + "]"
                k = src.find(close, lb[1])
                if k < 0:
                    raise LuaSyntaxError(f"line {line}: unterminated long comment")
                line += src.count("\n", i, k)
                i = k + len(close)
            else:
                k = src.find("\n", i)
                i = n if k < 0 else k
        elif c in "\"'":
            j = i + 1
            while True:
                if j >= n or src[j] == "\n":
                    raise LuaSyntaxError(f"line {line}: unterminated string")
                if src[j] == "\\":
                    j += 2
                    continue
                if src[j] == c:
                    break
                j += 1
            toks.append(("string", src[i + 1:j], line))
            i = j + 1
        elif c == "[" and _long_bracket(src, i):
            lvl, start = _long_bracket(src, i)
            close = "]" + "=" * lvl + "]"
            k = src.find(close, start)
            if k < 0:
                raise LuaSyntaxError(f"line {line}: unterminated long string")
            toks.append(("string", src[start:k], line))
            line += src.count("\n", i, k)
            i = k + len(close)
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = _NUM.match(src, i)
            toks.append(("number", m.group(), line))
            i = m.end()
        elif c.isalpha() or c == "_":
            m = _NAME.match(src, i)
            w = m.group()
            toks.append(("keyword" if w in KEYWORDS else "name", w, line))
            i = m.end()
        else:
            for op in _OPS:
                if src.startswith(op, i):
                    toks.append(("op", op, line))
                    i += len(op)
                    break
            else:
                raise LuaSyntaxError(f"line {line}: unexpected character {c!r}")
    return toks


def check_structure(toks):
    """block / bracket balance; raises LuaSyntaxError naming the line of the first offence"""
    stack = []          # entries: (kind, line)
    pairs = {")": "(", "}": "{", "]": "["}
    pending = []        # openers waiting for their `then` / `do`: 'if', 'elseif', 'for', 'while'
    for kind, v, line in toks:
        if kind == "op" and v in "({[":
            stack.append((v, line))
        elif kind == "op" and v in ")}]":
            if not stack or stack[-1][0] != pairs[v]:
                raise LuaSyntaxError(f"line {line}: unbalanced {v!r}")
            stack.pop()
        elif kind == "keyword":
            if v in ("function", "repeat"):
                stack.append((v, line))
            elif v == "if":
                stack.append(("if", line))
                pending.append(("then", len(stack)))
            elif v == "elseif":
                if not stack or stack[-1][0] != "if":
                    raise LuaSyntaxError(f"line {line}: 'elseif' outside an if block")
                pending.append(("then", len(stack)))
            elif v == "else":
                if not stack or stack[-1][0] != "if":
                    raise LuaSyntaxError(f"line {line}: 'else' outside an if block")
            elif v in ("for", "while"):
                stack.append((v, line))
                pending.append(("do", len(stack)))
            elif v == "then":
                if not pending or pending[-1] != ("then", len(stack)):
                    raise LuaSyntaxError(f"line {line}: 'then' without a matching if / elseif")
                pending.pop()
            elif v == "do":
                if pending and pending[-1] == ("do", len(stack)):
                    pending.pop()               # the `do` of a for / while: the block is already open
                else:
                    stack.append(("do", line))
            elif v == "end":
                if not stack or stack[-1][0] not in ("function", "if", "for", "while", "do"):
                    raise LuaSyntaxError(f"line {line}: 'end' closes nothing")
                if pending and pending[-1][1] == len(stack):
                    raise LuaSyntaxError(f"line {line}: block closed before its 'then' / 'do'")
                stack.pop()
            elif v == "until":
                if not stack or stack[-1][0] != "repeat":
                    raise LuaSyntaxError(f"line {line}: 'until' without 'repeat'")
                stack.pop()
    if stack:
        raise LuaSyntaxError(f"line {stack[-1][1]}: {stack[-1][0]!r} is never closed")
    if pending:
        raise LuaSyntaxError("an if / for / while has no 'then' / 'do'")
    return True


def lint(src):
    toks = tokenize(src)
    check_structure(toks)
    return toks


def defined_methods(toks, table):
    """names N of every `function table:N(` / `function table.N(` / `table.N = function` in the chunk"""
    out = set()
    for i, (k, v, _) in enumerate(toks):
        if k == "keyword" and v == "function" and i + 3 < len(toks):
            if toks[i + 1][1] == table and toks[i + 2][1] in (":", ".") and toks[i + 3][0] == "name":
                out.add(toks[i + 3][1])
        if k == "name" and v == table and i + 4 < len(toks) and toks[i + 1][1] == "." and toks[i + 2][0] == "name" \
                and toks[i + 3][1] == "=" and toks[i + 4][1] == "function":
            out.add(toks[i + 2][1])
    return out
