"""TLA+ lexer.

Token stream for the TLA+2 subset used by the spec corpus the reference ships
(/root/reference/*.tla, examples/**; lexeme inventory follows
examples/SpecifyingSystems/Syntax/TLAPlusGrammar.tla:7-66).  Every token keeps
its 1-based line and column because TLA+ conjunction/disjunction lists are
delimited by the *column* of their bullets.
"""
from __future__ import annotations

KEYWORDS = {
    "MODULE", "EXTENDS", "CONSTANT", "CONSTANTS", "VARIABLE", "VARIABLES",
    "ASSUME", "ASSUMPTION", "AXIOM", "THEOREM", "LEMMA", "PROPOSITION", "COROLLARY",
    "INSTANCE", "WITH", "LOCAL", "RECURSIVE",
    "IF", "THEN", "ELSE", "CASE", "OTHER", "LET", "IN", "CHOOSE", "LAMBDA",
    "EXCEPT", "ENABLED", "UNCHANGED", "SUBSET", "UNION", "DOMAIN",
    "TRUE", "FALSE", "BOOLEAN", "STRING",
    # proof keywords (proofs are skipped, but must lex)
    "BY", "DEF", "DEFS", "OBVIOUS", "OMITTED", "PROOF", "QED", "SUFFICES", "HAVE",
    "TAKE", "WITNESS", "PICK", "NEW", "DEFINE", "HIDE", "USE", "ONLY",
}

# backslash operators -> canonical name
BS_OPS = {
    "in": "\\in", "notin": "\\notin", "subseteq": "\\subseteq", "subset": "\\subset",
    "supseteq": "\\supseteq", "supset": "\\supset",
    "cup": "\\cup", "union": "\\cup", "cap": "\\cap", "intersect": "\\cap",
    "div": "\\div", "o": "\\o", "circ": "\\o", "X": "\\X", "times": "\\X",
    "leq": "<=", "geq": ">=", "land": "/\\", "lor": "\\/", "lnot": "~", "neg": "~",
    "equiv": "<=>", "A": "\\A", "E": "\\E", "AA": "\\AA", "EE": "\\EE",
    "cdot": "\\cdot", "oplus": "(+)", "ominus": "(-)", "otimes": "(\\X)", "oslash": "(/)",
    "odot": "(.)", "prec": "\\prec", "succ": "\\succ", "preceq": "\\preceq",
    "succeq": "\\succeq", "sim": "\\sim", "simeq": "\\simeq", "approx": "\\approx",
    "cong": "\\cong", "doteq": "\\doteq", "propto": "\\propto", "sqcap": "\\sqcap",
    "sqcup": "\\sqcup", "sqsubset": "\\sqsubset", "sqsupset": "\\sqsupset",
    "sqsubseteq": "\\sqsubseteq", "sqsupseteq": "\\sqsupseteq", "star": "\\star",
    "bullet": "\\bullet", "bigcirc": "\\bigcirc", "wr": "\\wr", "uplus": "\\uplus",
    "ll": "\\ll", "gg": "\\gg", "asymp": "\\asymp",
}

# multi-char symbolic operators, longest first
SYMS = [
    "<=>", "-+->", "|->", "...", "::=", "(\\X)",
    "==", "=>", "=<", "<=", ">=", "/=", "/\\", "\\/", "..", "<<", ">>", "<-", "->", "~>",
    "[]", "<>", "@@", ":>", "<:", "::", "||", "&&", "(+)", "(-)", "(.)", "(/)", "^+", "^*", "^#",
    "|-", "-|", "|=", "=|", "++", "--", "**", "//", "^^", "??", "%%", "##", "$$", "!!",
    "-.",      # name of the prefix minus operator (Standard/Integers.tla:6  -. a == 0 - a)
]
SINGLE = set("=#<>+-*/%^~()[]{},:;!'.@_|&$?")


class Tok:
    __slots__ = ("t", "v", "line", "col", "ecol")

    def __init__(self, t, v, line, col, ecol):
        self.t = t      # 'id','num','str','op','kw','sep','end','step','eof'
        self.v = v
        self.line = line
        self.col = col
        self.ecol = ecol  # column of last char

    def __repr__(self):
        return f"Tok({self.t},{self.v!r},{self.line}:{self.col})"


class LexError(Exception):
    pass


def _is_id_start(c):
    return c.isalpha() or c == "_"


def _is_id_char(c):
    return c.isalnum() or c == "_"


def lex(text: str, whole_file: bool = True):
    """Tokenise a module file.  With whole_file=True, text before the first
    `---- MODULE` line and after the closing `====` is ignored (TLA+ allows
    arbitrary prose there)."""
    toks = []
    i = 0
    n = len(text)
    line = 1
    lstart = 0  # index of current line start
    started = not whole_file
    depth_mod = 0

    def col_of(idx):
        # tabs count as advancing to next multiple of 8 -- approximate by 1 (specs use spaces)
        return idx - lstart + 1

    while i < n:
        c = text[i]
        if c == "\n":
            line += 1
            i += 1
            lstart = i
            continue
        if c in " \t\r\f":
            i += 1
            continue
        if not started:
            # look for ---- MODULE
            if c == "-" and text.startswith("----", i):
                j = i
                while j < n and text[j] == "-":
                    j += 1
                k = j
                while k < n and text[k] in " \t":
                    k += 1
                if text.startswith("MODULE", k):
                    started = True
                    toks.append(Tok("sep", "----", line, col_of(i), col_of(j - 1)))
                    i = j
                    continue
                i = j
                continue
            i += 1
            continue
        # comments
        if c == "(" and i + 1 < n and text[i + 1] == "*":
            depth = 1
            i += 2
            while i < n and depth > 0:
                if text.startswith("(*", i):
                    depth += 1
                    i += 2
                elif text.startswith("*)", i):
                    depth -= 1
                    i += 2
                else:
                    if text[i] == "\n":
                        line += 1
                        lstart = i + 1
                    i += 1
            continue
        if c == "\\" and i + 1 < n and text[i + 1] == "*":
            while i < n and text[i] != "\n":
                i += 1
            continue
        # separators
        if c == "-" and text.startswith("----", i):
            j = i
            while j < n and text[j] == "-":
                j += 1
            toks.append(Tok("sep", "----", line, col_of(i), col_of(j - 1)))
            i = j
            continue
        if c == "=" and text.startswith("====", i):
            j = i
            while j < n and text[j] == "=":
                j += 1
            toks.append(Tok("end", "====", line, col_of(i), col_of(j - 1)))
            i = j
            depth_mod = sum(1 for t in toks if t.t == "kw" and t.v == "MODULE") - \
                sum(1 for t in toks if t.t == "end")
            if whole_file and depth_mod <= 0:
                break
            continue
        # strings
        if c == '"':
            j = i + 1
            buf = []
            while j < n and text[j] != '"':
                if text[j] == "\\" and j + 1 < n:
                    e = text[j + 1]
                    buf.append({"n": "\n", "t": "\t", '"': '"', "\\": "\\", "r": "\r", "f": "\f"}.get(e, e))
                    j += 2
                else:
                    if text[j] == "\n":
                        raise LexError(f"unterminated string at line {line}")
                    buf.append(text[j])
                    j += 1
            toks.append(Tok("str", "".join(buf), line, col_of(i), col_of(j)))
            i = j + 1
            continue
        # numbers
        if c.isdigit():
            j = i
            while j < n and text[j].isdigit():
                j += 1
            # identifier starting with digits (e.g. 1a) is legal TLA+ but unused; treat as number
            if j < n and _is_id_start(text[j]) and text[j] != "_":
                k = j
                while k < n and _is_id_char(text[k]):
                    k += 1
                toks.append(Tok("id", text[i:k], line, col_of(i), col_of(k - 1)))
                i = k
                continue
            toks.append(Tok("num", int(text[i:j]), line, col_of(i), col_of(j - 1)))
            i = j
            continue
        # identifiers / keywords / WF_ SF_
        if _is_id_start(c) and not (c == "_" and not (i + 1 < n and _is_id_char(text[i + 1]))):
            j = i
            while j < n and _is_id_char(text[j]):
                j += 1
            w = text[i:j]
            if w in ("WF_", "SF_"):
                toks.append(Tok("op", w, line, col_of(i), col_of(j - 1)))
            elif (w.startswith("WF_") or w.startswith("SF_")) and len(w) > 3:
                # WF_vars  -> op WF_ + id vars
                toks.append(Tok("op", w[:3], line, col_of(i), col_of(i + 2)))
                toks.append(Tok("id", w[3:], line, col_of(i + 3), col_of(j - 1)))
            elif w in KEYWORDS:
                toks.append(Tok("kw", w, line, col_of(i), col_of(j - 1)))
            else:
                toks.append(Tok("id", w, line, col_of(i), col_of(j - 1)))
            i = j
            continue
        # backslash operators
        if c == "\\":
            if i + 1 < n and text[i + 1] == "/":
                toks.append(Tok("op", "\\/", line, col_of(i), col_of(i + 1)))
                i += 2
                continue
            j = i + 1
            while j < n and text[j].isalpha():
                j += 1
            w = text[i + 1:j]
            if w in ("b", "o", "h", "B", "O", "H") and j < n and text[j].isalnum() and w != "o":
                # based numbers \b101 \hFF
                k = j
                while k < n and text[k].isalnum():
                    k += 1
                base = {"b": 2, "o": 8, "h": 16}[w.lower()]
                toks.append(Tok("num", int(text[j:k], base), line, col_of(i), col_of(k - 1)))
                i = k
                continue
            if w == "":
                toks.append(Tok("op", "\\", line, col_of(i), col_of(i)))
                i += 1
                continue
            if w in BS_OPS:
                toks.append(Tok("op", BS_OPS[w], line, col_of(i), col_of(j - 1)))
                i = j
                continue
            raise LexError(f"unknown operator \\{w} at line {line}")
        # proof step tokens  <1>2.  <1>  <2>a.
        if c == "<" and i + 1 < n and text[i + 1].isdigit():
            j = i + 1
            while j < n and text[j].isdigit():
                j += 1
            if j < n and text[j] == ">":
                k = j + 1
                while k < n and (text[k].isalnum() or text[k] == "_"):
                    k += 1
                while k < n and text[k] == ".":
                    k += 1
                toks.append(Tok("step", text[i:k], line, col_of(i), col_of(k - 1)))
                i = k
                continue
        # ]_  and >>_
        if c == "]" and i + 1 < n and text[i + 1] == "_":
            toks.append(Tok("op", "]_", line, col_of(i), col_of(i + 1)))
            i += 2
            continue
        if c == ">" and text.startswith(">>_", i):
            toks.append(Tok("op", ">>_", line, col_of(i), col_of(i + 2)))
            i += 3
            continue
        matched = False
        for s in SYMS:
            if text.startswith(s, i):
                toks.append(Tok("op", s, line, col_of(i), col_of(i + len(s) - 1)))
                i += len(s)
                matched = True
                break
        if matched:
            continue
        if c in SINGLE:
            toks.append(Tok("op", c, line, col_of(i), col_of(i)))
            i += 1
            continue
        # tolerate stray non-ascii bytes (e.g. latin-1 quotes inside prose outside comments)
        if ord(c) > 127:
            i += 1
            continue
        raise LexError(f"unexpected character {c!r} at line {line} col {col_of(i)}")
    toks.append(Tok("eof", None, line, 1, 1))
    return toks
