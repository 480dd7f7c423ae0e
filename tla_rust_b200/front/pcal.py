"""PlusCal (p-syntax) -> TLA+ translator: the `pcal2tla` half of the reference's CLI
contract (Makefile:3-4, README.md:217-219,256-263).

Grammar and translation scheme follow examples/p-manual.pdf App. A (PDF p.60-62) and
App. B (p.63-67): one action per label, `pc`, `ProcSet`, `vars`, `Init`, `Next`, `Spec`,
`Termination`, the termination disjunct, process-local variables of a process *set*
turned into functions over the set, `assert e` -> Assert(e, "Failure of assertion at
line L, column C.").  The text layout (bullet columns, 78-column wrapping of
UNCHANGED lists, blank lines) mirrors the real translator so that the action
locations printed in error traces agree with README.md:278-316.

Supported statements: assignment (incl. `||` multi-assignment and `f[i] := e`), if /
elsif / else, while, either / or, with, await / when, assert, print, skip, goto.
Procedures, macros and `define` blocks with `call`/`return` are rejected with an error.
"""
from __future__ import annotations

import os
import re

from .lexer import lex, Tok

WRAP = 78


class PcalError(Exception):
    pass


class Stmt:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.label = None
        self.__dict__.update(kw)


KW_END_EXPR = {"then", "do", "or", "else", "elsif", "end", "begin", "process", "variables", "variable",
               "procedure", "macro", "define", "fair"}


def _tok_text(t: Tok):
    if t.t == "str":
        return '"' + t.v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if t.t == "num":
        return str(t.v)
    return str(t.v)


class PParser:
    def __init__(self, toks, line_off, col_off_first, macros=None):
        self.toks = toks
        self.p = 0
        self.line_off = line_off
        self.col_off_first = col_off_first
        self.macros = macros if macros is not None else {}     # name -> (parameter names, body tokens)

    def macro_def(self):
        """macro Name(p1, ..., pn) begin <statements> end macro [;]   (manual p.61): body kept as tokens."""
        self.expect_word("macro")
        name = self.next().v
        params = []
        self.expect_op("(")
        while not self.is_op(")"):
            params.append(self.next().v)
            if self.is_op(","):
                self.p += 1
        self.expect_op(")")
        self.expect_word("begin")
        body = []
        while not (self.is_word("end") and self.is_word("macro", 1)):
            t = self.next()
            if t.t == "eof":
                raise PcalError(f"PlusCal: macro {name} is not closed by `end macro`")
            body.append(t)
        self.p += 2
        self.skip_semi()
        self.macros[name] = (params, body)

    def macro_call(self, name):
        """Name(e1, ..., en): the body with every parameter token replaced by the (parenthesised) argument."""
        params, body = self.macros[name]
        self.p += 1
        self.expect_op("(")
        args = []
        while not self.is_op(")"):
            args.append(self.expr(stop_ops=(")",), stop_comma=True))
            if self.is_op(","):
                self.p += 1
        self.expect_op(")")
        if len(args) != len(params):
            raise PcalError(f"PlusCal: macro {name} expects {len(params)} arguments")
        sub = dict(zip(params, args))
        toks = []
        for t in body:
            if t.t == "id" and t.v in sub:
                a = sub[t.v]
                if len(a) == 1:
                    toks.append(a[0])
                else:
                    toks.append(Tok("op", "(", t.line, t.col, t.col))
                    toks.extend(a)
                    toks.append(Tok("op", ")", t.line, t.col, t.col))
            else:
                toks.append(t)
        toks.append(Tok("eof", None, body[-1].line if body else 0, 0, 0))
        inner = PParser(toks, self.line_off, self.col_off_first, self.macros)
        stmts = inner.stmt_seq(set())
        return Stmt("macro_expansion", body=stmts)

    def peek(self, k=0):
        i = self.p + k
        return self.toks[i] if i < len(self.toks) else self.toks[-1]

    def next(self):
        t = self.toks[self.p]
        self.p += 1
        return t

    def is_word(self, w, k=0):
        t = self.peek(k)
        return t.t in ("id", "kw") and t.v == w

    def is_op(self, v, k=0):
        t = self.peek(k)
        return t.t == "op" and t.v == v

    def expect_word(self, w):
        if not self.is_word(w):
            t = self.peek()
            raise PcalError(f"PlusCal: expected `{w}` at line {t.line + self.line_off}, found {t.v!r}")
        self.p += 1

    def expect_op(self, v):
        if not self.is_op(v):
            t = self.peek()
            raise PcalError(f"PlusCal: expected `{v}` at line {t.line + self.line_off}, found {t.v!r}")
        self.p += 1

    def skip_semi(self):
        while self.is_op(";"):
            self.p += 1

    def is_assign_op(self, k=0):
        a, b = self.peek(k), self.peek(k + 1)
        return a.t == "op" and a.v == ":" and b.t == "op" and b.v == "=" and a.line == b.line and b.col == a.col + 1

    def expr(self, stop_ops=(";",), stop_words=KW_END_EXPR, stop_comma=False):
        """Collect expression tokens until a terminator at bracket depth 0."""
        depth = 0
        out = []
        while True:
            t = self.peek()
            if t.t == "eof":
                break
            if t.t == "op":
                if depth == 0 and (t.v in stop_ops or (stop_comma and t.v == ",")):
                    break
                if depth == 0 and t.v == "||":
                    break
                if depth == 0 and self.is_assign_op():
                    break
                if t.v in ("(", "[", "{", "<<"):
                    depth += 1
                elif t.v in (")", "]", "}", ">>", "]_", ">>_"):
                    depth -= 1
            elif t.t in ("id", "kw") and depth == 0 and t.v in stop_words:
                # `or`/`end` etc. terminate; but IF/THEN/ELSE inside TLA exprs are upper-case keywords
                break
            out.append(self.next())
        if not out:
            t = self.peek()
            raise PcalError(f"PlusCal: expression expected at line {t.line + self.line_off}")
        return out

    # -- declarations ----------------------------------------------------------
    def var_decls(self):
        """variables x = e, y \\in S; z;  -> list of (name, kind, exprtoks)"""
        decls = []
        while True:
            t = self.peek()
            if t.t != "id" or t.v in ("process", "begin", "define", "macro", "procedure", "fair"):
                break
            name = self.next().v
            kind, e = None, None
            if self.is_op("="):
                self.p += 1
                kind = "="
                e = self.expr(stop_ops=(";",), stop_comma=True)
            elif self.is_op("\\in"):
                self.p += 1
                kind = "\\in"
                e = self.expr(stop_ops=(";",), stop_comma=True)
            decls.append((name, kind, e))
            if self.is_op(",") or self.is_op(";"):
                self.p += 1
                continue
            break
        return decls

    # -- statements ------------------------------------------------------------
    def stmt_seq(self, enders):
        stmts = []
        while True:
            self.skip_semi()
            t = self.peek()
            if t.t == "eof" or (t.t in ("id", "kw") and t.v in enders):
                break
            st = self.stmt()
            if st.kind == "macro_expansion":          # spliced in place; a label on the call labels its first statement
                if st.label and st.body:
                    st.body[0].label = st.label
                stmts.extend(st.body)
            else:
                stmts.append(st)
        return stmts

    def stmt(self):
        label = None
        t = self.peek()
        if t.t == "id" and self.is_op(":", 1) and not self.is_assign_op(1) and not self.is_op("::", 1):
            label = self.next().v
            self.p += 1
            if self.is_op("+") or self.is_op("-"):
                self.p += 1
        s = self.stmt_unlabeled()
        s.label = label
        return s

    def stmt_unlabeled(self):
        t = self.peek()
        if t.t == "id" and t.v in self.macros and self.is_op("(", 1):
            return self.macro_call(t.v)
        w = t.v if t.t in ("id", "kw") else None
        if w == "if":
            self.p += 1
            arms = []
            c = self.expr()
            self.expect_word("then")
            body = self.stmt_seq({"elsif", "else", "end"})
            arms.append((c, body))
            els = None
            while True:
                if self.is_word("elsif"):
                    self.p += 1
                    c = self.expr()
                    self.expect_word("then")
                    arms.append((c, self.stmt_seq({"elsif", "else", "end"})))
                    continue
                if self.is_word("else"):
                    self.p += 1
                    els = self.stmt_seq({"end"})
                break
            self.expect_word("end")
            self.expect_word("if")
            return Stmt("if", arms=arms, els=els)
        if w == "while":
            self.p += 1
            c = self.expr()
            self.expect_word("do")
            body = self.stmt_seq({"end"})
            self.expect_word("end")
            self.expect_word("while")
            return Stmt("while", cond=c, body=body)
        if w == "either":
            self.p += 1
            alts = [self.stmt_seq({"or", "end"})]
            while self.is_word("or"):
                self.p += 1
                alts.append(self.stmt_seq({"or", "end"}))
            self.expect_word("end")
            self.expect_word("either")
            return Stmt("either", alts=alts)
        if w == "with":
            self.p += 1
            binds = []
            while True:
                v = self.next().v
                if self.is_op("="):
                    k = "="
                elif self.is_op("\\in"):
                    k = "\\in"
                else:
                    raise PcalError("PlusCal: bad with-binding")
                self.p += 1
                e = self.expr(stop_ops=(";", ","), stop_words={"do"})
                binds.append((v, k, e))
                if self.is_op(",") or self.is_op(";"):
                    self.p += 1
                    continue
                break
            self.expect_word("do")
            body = self.stmt_seq({"end"})
            self.expect_word("end")
            self.expect_word("with")
            return Stmt("with", binds=binds, body=body)
        if w in ("await", "when"):
            self.p += 1
            return Stmt("await", cond=self.expr())
        if w == "assert":
            self.p += 1
            line = t.line + self.line_off
            col = t.col + (self.col_off_first if t.line == 1 else 0)
            return Stmt("assert", cond=self.expr(), line=line, col=col)
        if w == "print":
            self.p += 1
            return Stmt("print", e=self.expr())
        if w == "skip":
            self.p += 1
            return Stmt("skip")
        if w == "goto":
            self.p += 1
            return Stmt("goto", target=self.next().v)
        if w in ("call", "return"):
            raise PcalError("PlusCal procedures (call/return) are not supported")
        # assignment(s)
        assigns = []
        while True:
            lhs = self.next()
            if lhs.t != "id":
                raise PcalError(f"PlusCal: statement expected at line {lhs.line + self.line_off}, found {lhs.v!r}")
            idx = []
            while self.is_op("[") or self.is_op("."):
                if self.is_op("."):
                    self.p += 1
                    idx.append(("fld", self.next().v))
                else:
                    self.p += 1
                    depth = 1
                    toks = []
                    while True:
                        x = self.next()
                        if x.t == "op" and x.v in ("[", "(", "{", "<<"):
                            depth += 1
                        if x.t == "op" and x.v in ("]", ")", "}", ">>"):
                            depth -= 1
                            if depth == 0:
                                break
                        toks.append(x)
                    idx.append(("idx", toks))
            if not self.is_assign_op():
                raise PcalError(f"PlusCal: `:=` expected at line {lhs.line + self.line_off}")
            self.p += 2
            rhs = self.expr()
            assigns.append((lhs.v, idx, rhs))
            if self.is_op("||"):
                self.p += 1
                continue
            break
        return Stmt("assign", assigns=assigns)


class Proc:
    def __init__(self, name, kind, idexpr, decls, body):
        self.name = name
        self.kind = kind      # '=' single process, '\\in' process set, None uniprocess
        self.idexpr = idexpr
        self.decls = decls
        self.body = body


# --------------------------------------------------------------------------------
class Out:
    """Conjunction tree -> text."""

    @staticmethod
    def conj(items, indent):
        """items: list of str (may be multi-line, continuation lines already absolute) or nested tuples."""
        lines = []
        pad = " " * indent
        for it in items:
            sub = Out.render(it, indent + 3)
            lines.append(pad + "/\\ " + sub[0].lstrip())
            lines.extend(sub[1:])
        return lines

    @staticmethod
    def render(it, indent):
        """Return list of lines; the first line is to be placed at column `indent` (0-based)."""
        pad = " " * indent
        if isinstance(it, str):
            parts = it.split("\n")
            return [pad + parts[0]] + parts[1:]
        k = it[0]
        if k == "conj":
            return Out.conj(it[1], indent)
        if k == "if":
            _, cond, th, el = it
            lines = [pad + "IF " + cond]
            tl = Out.conj(th, indent + 3 + 5)
            tl[0] = " " * (indent + 3) + "THEN " + tl[0].lstrip()
            el_l = Out.conj(el, indent + 3 + 5)
            el_l[0] = " " * (indent + 3) + "ELSE " + el_l[0].lstrip()
            return lines + tl + el_l
        if k == "or":
            lines = []
            for alt in it[1]:
                al = Out.conj(alt, indent + 3)
                al[0] = pad + "\\/ " + al[0].lstrip()
                lines.extend(al)
            return lines
        if k == "exists":
            _, hdr, body = it
            return [pad + hdr] + Out.conj(body, indent + 2)
        if k == "let":
            _, hdr, body = it
            return [pad + hdr] + Out.conj(body, indent + 2)
        raise PcalError("internal: bad output node")


def wrap_list(prefix, names, closing=" >>", first_indent=None):
    """`prefix` already contains '<< '; wrap at WRAP columns aligning under the first item."""
    lines = []
    cur = prefix
    align = " " * len(prefix) if first_indent is None else first_indent
    for i, nm in enumerate(names):
        piece = nm + (", " if i + 1 < len(names) else closing)
        if len(cur) + len(piece) > WRAP and cur.strip() and cur != prefix:
            lines.append(cur)
            cur = align + piece
        else:
            cur += piece
    lines.append(cur)
    return "\n".join(lines)


class Translator:
    def __init__(self, text, line_off=0, col_off_first=0):
        self.text = text
        self.define_text = None
        self.toks = lex(text, whole_file=False)
        self.pp = PParser(self.toks, line_off, col_off_first)
        self.line_off = line_off

    # -- parse ------------------------------------------------------------------
    def parse(self):
        pp = self.pp
        t = pp.next()
        if not (t.t == "op" and t.v == "--"):
            raise PcalError("PlusCal: expected --algorithm")
        if pp.is_word("fair"):
            pp.p += 1
        pp.expect_word("algorithm")
        self.name = pp.next().v
        self.gdecls = []
        if pp.is_word("variables") or pp.is_word("variable"):
            pp.p += 1
            self.gdecls = pp.var_decls()
        if pp.is_word("define"):
            # define <TLA+ definitions> end define [;]  (manual p.61): copied verbatim into the translation, after the
            # declaration of the global variables (the definitions may mention them)
            pp.p += 1
            first = pp.peek()
            last = None
            while not (pp.is_word("end") and pp.is_word("define", 1)):
                last = pp.next()
                if last.t == "eof":
                    raise PcalError("PlusCal: `define` is not closed by `end define`")
            pp.p += 2
            pp.skip_semi()
            lines = self.text.split("\n")
            if last is not None:
                seg = lines[first.line - 1:last.line]
                seg[-1] = seg[-1][:last.ecol]
                seg[0] = " " * (first.col - 1) + seg[0][first.col - 1:]
                ind = min(len(x) - len(x.lstrip()) for x in seg if x.strip())
                self.define_text = "\n".join(x[ind:].rstrip() for x in seg)
        while pp.is_word("macro"):
            pp.macro_def()
        if pp.is_word("procedure"):
            raise PcalError("PlusCal procedures are not supported")
        self.procs = []
        if pp.is_word("begin"):
            pp.p += 1
            body = pp.stmt_seq({"end"})
            pp.expect_word("end")
            pp.expect_word("algorithm")
            self.procs.append(Proc(None, None, None, [], body))
            return
        while pp.is_word("process") or pp.is_word("fair"):
            if pp.is_word("fair"):
                pp.p += 1
                if pp.is_op("+"):
                    pp.p += 1
            pp.expect_word("process")
            name = pp.next().v
            if pp.is_op("="):
                kind = "="
            elif pp.is_op("\\in"):
                kind = "\\in"
            else:
                raise PcalError("PlusCal: process needs `=` or `\\in`")
            pp.p += 1
            idexpr = pp.expr(stop_ops=(";",), stop_words={"variables", "variable", "begin"})
            decls = []
            if pp.is_word("variables") or pp.is_word("variable"):
                pp.p += 1
                decls = pp.var_decls()
            pp.expect_word("begin")
            body = pp.stmt_seq({"end"})
            pp.expect_word("end")
            pp.expect_word("process")
            pp.skip_semi()
            self.procs.append(Proc(name, kind, idexpr, decls, body))
        pp.expect_word("end")
        pp.expect_word("algorithm")

    # -- expression text --------------------------------------------------------
    def etext(self, toks, primed=frozenset(), selfvars=frozenset(), bound=frozenset()):
        out = []
        prev = None
        for i, t in enumerate(toks):
            s = _tok_text(t)
            if t.t == "id" and t.v not in bound:
                prv = toks[i - 1] if i > 0 else None
                nxt = toks[i + 1] if i + 1 < len(toks) else None
                is_field = (prv is not None and prv.t == "op" and prv.v == ".") or \
                           (nxt is not None and nxt.t == "op" and nxt.v == "|->")
                if not is_field:
                    if t.v in primed:
                        s += "'"
                    if t.v in selfvars:
                        s += "[self]"
            if prev is not None:
                if t.line == prev.line:
                    gap = t.col - prev.ecol - 1
                    out.append(" " * max(gap, 0))
                else:
                    out.append(" ")
            out.append(s)
            prev = t
        return "".join(out)

    # -- translation ------------------------------------------------------------
    def translate(self):
        self.parse()
        uni = self.procs[0].kind is None
        self.uni = uni
        gvars = [d[0] for d in self.gdecls]
        lvars = []
        for p in self.procs:
            lvars += [d[0] for d in p.decls]
        self.allvars = gvars + ["pc"] + lvars
        out = []
        out.append("VARIABLES " + wrap_list("", self.allvars, closing="").replace("\n", "\n          "))
        out.append("")
        if self.define_text:
            out.append("(* define statement *)")
            out.extend(self.define_text.split("\n"))
            out.append("")
        out.append(wrap_list("vars == << ", self.allvars))
        out.append("")
        if not uni:
            sets = []
            for p in self.procs:
                e = self.etext(p.idexpr)
                sets.append("(" + e + ")" if p.kind == "\\in" else "{" + e + "}")
            out.append("ProcSet == " + " \\cup ".join(sets))
            out.append("")
        # Init
        init = []
        if self.gdecls:
            init.append("(* Global variables *)")
            for nm, k, e in self.gdecls:
                init.append(f"/\\ {nm} {'=' if k != chr(92) + 'in' else chr(92) + 'in'} " +
                            (self.etext(e) if e else "defaultInitValue"))
        for p in self.procs:
            if p.decls:
                init.append(f"(* Process {p.name} *)")
                for nm, k, e in p.decls:
                    et = self.etext(e) if e else "defaultInitValue"
                    if p.kind == "\\in":
                        pe = self.etext(p.idexpr)
                        if k == "\\in":
                            init.append(f"/\\ {nm} \\in [{pe} -> {et}]")
                        else:
                            init.append(f"/\\ {nm} = [self \\in {pe} |-> {et}]")
                    else:
                        init.append(f"/\\ {nm} {k or '='} {et}")
        first_labels = []
        for p in self.procs:
            if not p.body or p.body[0].label is None:
                raise PcalError(f"PlusCal: first statement of process {p.name} must be labeled")
            first_labels.append(p.body[0].label)
        if uni:
            init.append(f'/\\ pc = "{first_labels[0]}"')
        elif len(self.procs) == 1:
            init.append(f'/\\ pc = [self \\in ProcSet |-> "{first_labels[0]}"]')
        else:
            arms = []
            for p, fl in zip(self.procs, first_labels):
                e = self.etext(p.idexpr)
                arms.append((f"self \\in {e}" if p.kind == "\\in" else f"self = {e}") + f' -> "{fl}"')
            pre = "/\\ pc = [self \\in ProcSet |-> CASE "
            init.append(pre + arms[0] + "".join("\n" + " " * (8 + len(pre) - 5) + "[] " + a for a in arms[1:]) + "]")
        out.append("Init == " + init[0])
        for ln in init[1:]:
            parts = ln.split("\n")
            out.append(" " * 8 + parts[0])
            out.extend(parts[1:])
        out.append("")
        # actions
        self.actions = []   # (procindex, label, text lines)
        proc_action_names = []
        for p in self.procs:
            self.cur = p
            self.selfvars = frozenset(d[0] for d in p.decls) if p.kind == "\\in" else frozenset()
            self.labels_of_proc = []
            self.pending = []
            self.gen_actions(p)
            proc_action_names.append(list(self.labels_of_proc))
        for txt in self.actions:
            out.extend(txt)
            out.append("")
            if txt and txt[0].startswith("\x00"):
                pass
        # Next
        singles, sets = [], []
        for p, labs in zip(self.procs, proc_action_names):
            if uni:
                continue
            if p.kind == "\\in":
                sets.append(f"(\\E self \\in {self.etext(p.idexpr)}: {p.name}(self))")
            else:
                singles.append(p.name)
        if uni:
            disj = list(proc_action_names[0])
            term = '(pc = "Done" /\\ UNCHANGED vars)'
        else:
            disj = singles + sets
            term = '((\\A self \\in ProcSet: pc[self] = "Done") /\\ UNCHANGED vars)'
        nxt = ["Next == " + disj[0]]
        for dj in disj[1:]:
            nxt.append(" " * 11 + "\\/ " + dj)
        nxt.append(" " * 11 + "\\/ (* Disjunct to prevent deadlock on termination *)")
        nxt.append(" " * 14 + term)
        out.extend(nxt)
        out.append("")
        out.append("Spec == Init /\\ [][Next]_vars")
        out.append("")
        if uni:
            out.append('Termination == <>(pc = "Done")')
        else:
            out.append('Termination == <>(\\A self \\in ProcSet: pc[self] = "Done")')
        out.append("")
        return out

    def pc_read(self):
        if self.uni:
            return "pc"
        if self.cur.kind == "\\in":
            return "pc[self]"
        return f"pc[{self.etext(self.cur.idexpr)}]"

    def pc_set(self, label):
        if self.uni:
            return f"pc' = \"{label}\""
        if self.cur.kind == "\\in":
            return f"pc' = [pc EXCEPT ![self] = \"{label}\"]"
        return f"pc' = [pc EXCEPT ![{self.etext(self.cur.idexpr)}] = \"{label}\"]"

    def gen_actions(self, p):
        # worklist of (label, stmts-from-label, continuation label)
        self.done_labels = set()
        self.work = [(p.body[0].label, p.body, 0, "Done")]
        # collect labels in textual order for the process disjunction
        order = []
        self._collect_labels(p.body, order)
        produced = {}
        while self.work:
            label, seq, i, cont = self.work.pop(0)
            if label in self.done_labels:
                continue
            self.done_labels.add(label)
            produced[label] = self.gen_action(label, seq, i, cont)
        for lb in order:
            if lb in produced:
                self.actions.append(produced[lb])
                self.labels_of_proc.append(lb)
        if not self.uni:
            arg = "(self)" if p.kind == "\\in" else ""
            names = [lb + arg for lb in self.labels_of_proc]
            self.actions.append([wrap_list(f"{p.name}{arg} == ", names, closing="").replace(", ", " \\/ ")])

    def _collect_labels(self, stmts, out):
        for s in stmts:
            if s.label:
                out.append(s.label)
            if s.kind == "if":
                for _, b in s.arms:
                    self._collect_labels(b, out)
                if s.els:
                    self._collect_labels(s.els, out)
            elif s.kind == "while":
                self._collect_labels(s.body, out)
            elif s.kind == "either":
                for b in s.alts:
                    self._collect_labels(b, out)
            elif s.kind == "with":
                self._collect_labels(s.body, out)

    @staticmethod
    def has_label(stmts):
        for s in stmts:
            if s.label:
                return True
            if s.kind == "if":
                if any(Translator.has_label(b) for _, b in s.arms) or (s.els and Translator.has_label(s.els)):
                    return True
            elif s.kind == "while" and Translator.has_label(s.body):
                return True
            elif s.kind == "either" and any(Translator.has_label(b) for b in s.alts):
                return True
            elif s.kind == "with" and Translator.has_label(s.body):
                return True
        return False

    def gen_action(self, label, seq, i, cont):
        arg = "(self)" if (not self.uni and self.cur.kind == "\\in") else ""
        head = f"{label}{arg} == "
        first = seq[i]
        items = [f'{self.pc_read()} = "{label}"']
        assigned = set()
        if first.kind == "while":
            # L: while c do body end while; rest
            body_items, body_asg = self.gen_seq(first.body, 0, label, set(), bound=frozenset())
            rest_items, rest_asg = self.gen_seq(seq, i + 1, cont, set(), bound=frozenset())
            allasg = body_asg | rest_asg
            self._pad_unchanged(body_items, body_asg, allasg)
            self._pad_unchanged(rest_items, rest_asg, allasg)
            items.append(("if", self.etext(first.cond, selfvars=self.selfvars), body_items, rest_items))
            assigned = allasg
        else:
            its, assigned = self.gen_seq(seq, i, cont, set(), bound=frozenset(), first=True)
            items.extend(its)
        unch = [v for v in self.allvars if v not in assigned and v != "pc"]
        if "pc" not in assigned:
            unch = [v for v in self.allvars if v not in assigned]
        lines = Out.conj(items, len(head))
        if unch:
            pre = " " * len(head) + "/\\ UNCHANGED "
            if len(unch) == 1:
                lines.append(pre + unch[0])
            else:
                lines.extend(wrap_list(pre + "<< ", unch).split("\n"))
        lines[0] = head + lines[0].lstrip()
        return lines

    def _pad_unchanged(self, items, asg, allasg):
        missing = [v for v in self.allvars if v in allasg and v not in asg]
        if missing:
            if not [x for x in items if not (isinstance(x, str) and x.startswith("pc' ="))] and False:
                pass
            if len(missing) == 1:
                items.append("UNCHANGED " + missing[0])
            else:
                items.append("UNCHANGED << " + ", ".join(missing) + " >>")

    def gen_seq(self, seq, i, cont, assigned, bound, first=False):
        """Translate seq[i:] into conjunct items for the current action.
        Returns (items, assigned-vars incl. 'pc')."""
        items = []
        assigned = set(assigned)
        while i < len(seq):
            s = seq[i]
            if s.label is not None and not first:
                # control passes to a new action
                items.append(self.pc_set(s.label))
                assigned.add("pc")
                self.work.append((s.label, seq, i, cont))
                return items, assigned
            first = False
            k = s.kind
            primed = frozenset(assigned - {"pc"})
            sv = self.selfvars
            if k == "assign":
                newly = []
                for var, idx, rhs in s.assigns:
                    rt = self.etext(rhs, primed, sv, bound)
                    path = ""
                    if var in sv:
                        path += "![self]"
                    for kind, ix in idx:
                        if kind == "fld":
                            path = (path or "!") + "." + ix
                        else:
                            path = (path or "!") + "[" + self.etext(ix, primed, sv, bound) + "]"
                    if path:
                        if not path.startswith("!"):
                            path = "!" + path
                        base = var + ("'" if var in assigned else "")
                        items.append(f"{var}' = [{base} EXCEPT {path} = {rt}]")
                    else:
                        items.append(f"{var}' = {rt}")
                    newly.append(var)
                assigned.update(newly)
            elif k == "await":
                items.append(self.etext(s.cond, primed, sv, bound))
            elif k == "assert":
                c = self.etext(s.cond, primed, sv, bound)
                items.append(f"Assert({c}, \n" + "\x01" +
                             f'"Failure of assertion at line {s.line}, column {s.col}.")')
            elif k == "print":
                items.append(f"PrintT({self.etext(s.e, primed, sv, bound)})")
            elif k == "skip":
                items.append("TRUE")
            elif k == "goto":
                items.append(self.pc_set(s.target))
                assigned.add("pc")
                return items, assigned
            elif k == "if":
                labeled = any(self.has_label(b) for _, b in s.arms) or bool(s.els and self.has_label(s.els))
                nxt_label = None
                if labeled:
                    # branches jump; continuation must be a labeled statement (or end of process)
                    if i + 1 < len(seq):
                        if seq[i + 1].label is None:
                            raise PcalError("PlusCal: statement after an `if` containing labels must be labeled")
                        nxt_label = seq[i + 1].label
                        self.work.append((nxt_label, seq, i + 1, cont))
                    else:
                        nxt_label = cont
                    branches = []
                    for c, b in s.arms:
                        bi, ba = self.gen_seq(b, 0, nxt_label, assigned, bound)
                        branches.append((self.etext(c, primed, sv, bound), bi, ba))
                    ei, ea = self.gen_seq(s.els or [], 0, nxt_label, assigned, bound)
                    allasg = set(ea)
                    for _, _, ba in branches:
                        allasg |= ba
                    for _, bi, ba in branches:
                        self._pad_unchanged(bi, ba, allasg)
                    self._pad_unchanged(ei, ea, allasg)
                    node = None
                    for c, bi, _ in reversed(branches):
                        node = ("if", c, bi, ei if node is None else [node])
                    items.append(node)
                    return items, allasg
                branches = []
                for c, b in s.arms:
                    bi, ba = self.gen_flat(b, assigned, bound)
                    branches.append((self.etext(c, primed, sv, bound), bi, ba))
                ei, ea = self.gen_flat(s.els or [], assigned, bound)
                allasg = set(ea)
                for _, _, ba in branches:
                    allasg |= ba
                for _, bi, ba in branches:
                    if not bi:
                        bi.append("TRUE")
                    self._pad_unchanged(bi, ba, allasg)
                if not ei:
                    ei.append("TRUE")
                self._pad_unchanged(ei, ea, allasg)
                node = None
                for c, bi, _ in reversed(branches):
                    node = ("if", c, bi, ei if node is None else [node])
                items.append(node)
                assigned = allasg
            elif k == "either":
                if any(self.has_label(b) for b in s.alts):
                    raise PcalError("PlusCal: labels inside `either` are not supported")
                alts = []
                allasg = set(assigned)
                for b in s.alts:
                    bi, ba = self.gen_flat(b, assigned, bound)
                    alts.append((bi, ba))
                    allasg |= ba
                for bi, ba in alts:
                    if not bi:
                        bi.append("TRUE")
                    self._pad_unchanged(bi, ba, allasg)
                items.append(("or", [bi for bi, _ in alts]))
                assigned = allasg
            elif k == "with":
                if self.has_label(s.body):
                    raise PcalError("PlusCal: labels inside `with` are not allowed")
                b2 = set(bound)
                hdrs = []
                for v, kk, e in s.binds:
                    et = self.etext(e, primed, sv, frozenset(b2))
                    hdrs.append((v, kk, et))
                    b2.add(v)
                bi, ba = self.gen_flat(s.body, assigned, frozenset(b2))
                node = bi
                for v, kk, et in reversed(hdrs):
                    if kk == "\\in":
                        node = [("exists", f"\\E {v} \\in {et}:", node)]
                    else:
                        node = [("let", f"LET {v} == {et} IN", node)]
                items.extend(node)
                assigned = ba
            elif k == "while":
                raise PcalError("PlusCal: `while` must be labeled")
            else:
                raise PcalError(f"PlusCal: unsupported statement {k}")
            i += 1
        items.append(self.pc_set(cont))
        assigned.add("pc")
        return items, assigned

    def gen_flat(self, stmts, assigned, bound):
        """Label-free statement list inside if/either/with: no pc handling."""
        save_work = self.work
        items = []
        asg = set(assigned)
        for s in stmts:
            sub, asg2 = self.gen_seq([s], 0, "\x00", asg, bound, first=True)
            # drop the pc' = continuation conjunct gen_seq appends
            sub = [x for x in sub if not (isinstance(x, str) and "\x00" in x)]
            asg2.discard("pc") if "pc" not in asg else None
            items.extend(sub)
            asg = asg2
        self.work = save_work
        return items, asg


def _finish_lines(lines):
    """Resolve the Assert continuation marker: align the message under the first argument."""
    out = []
    for ln in lines:
        if "\x01" in ln:
            prev = out[-1]
            col = prev.index("Assert(") + len("Assert(")
            out.append(" " * col + ln.replace("\x01", "").lstrip())
        else:
            out.append(ln)
    return out


ALG_RE = re.compile(r"--(fair\s+)?algorithm")


def translate_text(src: str):
    """Return (new_module_text, had_algorithm)."""
    m = ALG_RE.search(src)
    if not m:
        return src, False
    start = m.start()
    # the algorithm sits inside a (* ... *) comment: find its end (balanced)
    cstart = src.rfind("(*", 0, start)
    depth = 0
    i = cstart
    end = None
    while i < len(src):
        if src.startswith("(*", i):
            depth += 1
            i += 2
        elif src.startswith("*)", i):
            depth -= 1
            i += 2
            if depth == 0:
                end = i
                break
        else:
            i += 1
    if end is None:
        raise PcalError("unterminated comment around the algorithm")
    alg_text = src[start:end - 2]
    line_off = src.count("\n", 0, start)
    col_off = start - (src.rfind("\n", 0, start) + 1)
    tr = Translator(alg_text, line_off, col_off)
    body = _finish_lines([x for ln in tr.translate() for x in ln.split("\n")])
    block = ["\\* BEGIN TRANSLATION"] + body + ["\\* END TRANSLATION"]
    # existing translation?
    b = src.find("\\* BEGIN TRANSLATION")
    if b >= 0:
        e = src.find("\\* END TRANSLATION", b)
        if e < 0:
            raise PcalError("BEGIN TRANSLATION without END TRANSLATION")
        eol = src.find("\n", e)
        eol = len(src) if eol < 0 else eol
        new = src[:b] + "\n".join(block) + src[eol:]
        return new, True
    # insert after the line holding the end of the comment
    eol = src.find("\n", end)
    if eol < 0:
        eol = len(src)
    new = src[:eol + 1] + "\n".join(block) + "\n" + src[eol + 1:]
    return new, True


def translate_file(path: str, write_cfg=True, backup=True):
    """In-place translation with FILE.old backup and FILE.cfg creation (manual p.70-72)."""
    with open(path, "rb") as f:
        raw = f.read()
    try:
        src = raw.decode("utf-8")
    except UnicodeDecodeError:
        src = raw.decode("latin-1")
    new, had = translate_text(src)
    if not had:
        return False
    if backup:
        with open(os.path.splitext(path)[0] + ".old", "w") as f:
            f.write(src)
    with open(path, "w") as f:
        f.write(new)
    cfgp = os.path.splitext(path)[0] + ".cfg"
    if write_cfg and not os.path.exists(cfgp):
        with open(cfgp, "w") as f:
            f.write("SPECIFICATION Spec\n\\* Add statements after this line.\n")
    return True
