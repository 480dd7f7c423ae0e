"""TLA+ parser (hand-written Pratt parser with column-sensitive junction lists).

Grammar reference: examples/SpecifyingSystems/Syntax/TLAPlusGrammar.tla:68-254
(TLA+ v1) plus the TLA+2 constructs the corpus uses: RECURSIVE, LAMBDA,
`Def!n` / `Thm!:` selectors, named ASSUME/THEOREM, proofs (skipped).

AST: Node(k, a, line, col, eline, ecol) -- `k` is the kind, `a` a tuple of
children / payload.  Kinds are listed in `KINDS` below.
"""
from __future__ import annotations

from .lexer import lex, Tok


class ParseError(Exception):
    pass


class Node:
    __slots__ = ("k", "a", "line", "col", "eline", "ecol")

    def __init__(self, k, a, line=0, col=0, eline=0, ecol=0):
        self.k = k
        self.a = a
        self.line = line
        self.col = col
        self.eline = eline
        self.ecol = ecol

    def __repr__(self):
        return f"{self.k}{self.a!r}"

    def loc(self):
        return (self.line, self.col, self.eline, self.ecol)


class OpDef:
    """Operator / function definition.  params: list of (name, arity)."""
    __slots__ = ("name", "params", "body", "local", "module", "recursive", "line", "col", "eline", "ecol")

    def __init__(self, name, params, body, local=False, module=None):
        self.name = name
        self.params = params
        self.body = body
        self.local = local
        self.module = module
        self.recursive = False
        self.line = self.col = self.eline = self.ecol = 0

    def __repr__(self):
        return f"OpDef({self.name}/{len(self.params)})"


class Instance:
    __slots__ = ("name", "params", "module", "substs", "local")

    def __init__(self, name, params, module, substs, local=False):
        self.name = name      # None for unnamed INSTANCE
        self.params = params
        self.module = module
        self.substs = substs  # list of (ident, expr)
        self.local = local


class Module:
    def __init__(self, name):
        self.name = name
        self.extends = []
        self.constants = []   # (name, arity)
        self.variables = []
        self.defs = {}        # name -> OpDef  (insertion ordered)
        self.assumes = []     # (name|None, expr)
        self.theorems = []    # (name|None, expr)
        self.instances = []   # Instance
        self.submodules = {}


# infix operator table: name -> (lo, hi, left_assoc)
INFIX = {
    "=>": (1, 1, False), "<=>": (2, 2, False), "~>": (2, 2, False), "-+->": (2, 2, False),
    "/\\": (3, 3, True), "\\/": (3, 3, True),
    "=": (5, 5, False), "#": (5, 5, False), "/=": (5, 5, False), "<": (5, 5, False), ">": (5, 5, False),
    "<=": (5, 5, False), "=<": (5, 5, False), ">=": (5, 5, False), "\\in": (5, 5, False),
    "\\notin": (5, 5, False), "\\subseteq": (5, 5, False), "\\subset": (5, 5, False),
    "\\supseteq": (5, 5, False), "\\supset": (5, 5, False),
    "\\prec": (5, 5, False), "\\succ": (5, 5, False), "\\preceq": (5, 5, False), "\\succeq": (5, 5, False),
    "\\sim": (5, 5, False), "\\simeq": (5, 5, False), "\\approx": (5, 5, False), "\\cong": (5, 5, False),
    "\\doteq": (5, 5, False), "\\propto": (5, 5, False), "\\sqsubset": (5, 5, False),
    "\\sqsupset": (5, 5, False), "\\sqsubseteq": (5, 5, False), "\\sqsupseteq": (5, 5, False),
    "\\ll": (5, 5, False), "\\gg": (5, 5, False), "\\asymp": (5, 5, False),
    "::=": (5, 5, False), ":=": (5, 5, False), "|-": (5, 5, False), "-|": (5, 5, False),
    "|=": (5, 5, False), "=|": (5, 5, False),
    "\\cdot": (5, 14, True),
    "@@": (6, 6, True), ":>": (7, 7, False), "<:": (7, 7, False),
    "\\": (8, 8, False), "\\cup": (8, 8, True), "\\cap": (8, 8, True),
    "..": (9, 9, False), "...": (9, 9, False),
    "!!": (9, 13, False), "##": (9, 13, True), "$$": (9, 13, True), "$": (9, 13, True), "??": (9, 13, True),
    "\\sqcap": (9, 13, True), "\\sqcup": (9, 13, True), "\\uplus": (9, 13, True), "\\wr": (9, 14, False),
    "(+)": (10, 10, True), "+": (10, 10, True), "++": (10, 10, True), "%": (10, 11, False),
    "%%": (10, 11, True), "|": (10, 11, True), "||": (10, 11, True),
    "(-)": (11, 11, True), "-": (11, 11, True), "--": (11, 11, True),
    "&": (13, 13, True), "&&": (13, 13, True), "(.)": (13, 13, True), "(/)": (13, 13, False),
    "(\\X)": (13, 13, True), "*": (13, 13, True), "**": (13, 13, True), "/": (13, 13, False),
    "//": (13, 13, False), "\\bigcirc": (13, 13, True), "\\bullet": (13, 13, True),
    "\\div": (13, 13, False), "\\o": (13, 13, True), "\\star": (13, 13, True),
    "\\X": (10, 13, True),
    "^": (14, 14, False), "^^": (14, 14, False),
}
POSTFIX = {"^+", "^*", "^#"}

UNIT_KW = {"EXTENDS", "CONSTANT", "CONSTANTS", "VARIABLE", "VARIABLES", "ASSUME", "ASSUMPTION",
           "AXIOM", "THEOREM", "LEMMA", "PROPOSITION", "COROLLARY", "INSTANCE", "LOCAL", "RECURSIVE"}


class Parser:
    def __init__(self, toks):
        self.toks = toks
        self.p = 0
        self.jstack = []  # columns of enclosing junction lists

    # -- token access ------------------------------------------------------
    def raw(self, k=0):
        i = self.p + k
        return self.toks[i] if i < len(self.toks) else self.toks[-1]

    def peek(self, k=0):
        t = self.raw(k)
        if self.jstack and t.t != "eof" and t.col <= self.jstack[-1]:
            return _CUT
        return t

    def next(self):
        t = self.peek()
        if t is _CUT:
            raise self.err("unexpected end of junction item")
        self.p += 1
        return t

    def err(self, msg):
        t = self.raw()
        return ParseError(f"{msg} at line {t.line} col {t.col} (token {t.v!r})")

    def is_op(self, v, k=0):
        t = self.peek(k)
        return t.t == "op" and t.v == v

    def is_kw(self, v, k=0):
        t = self.peek(k)
        return t.t == "kw" and t.v == v

    def expect_op(self, v):
        t = self.peek()
        if t.t == "op" and t.v == v:
            self.p += 1
            return t
        raise self.err(f"expected {v!r}")

    def expect_kw(self, v):
        t = self.peek()
        if t.t == "kw" and t.v == v:
            self.p += 1
            return t
        raise self.err(f"expected {v}")

    def expect_id(self):
        t = self.peek()
        if t.t == "id":
            self.p += 1
            return t
        raise self.err("expected identifier")

    def mk(self, k, a, start, end=None):
        if end is None:
            end = self.toks[self.p - 1]
        if isinstance(start, Node):
            l, c = start.line, start.col
        else:
            l, c = start.line, start.col
        return Node(k, a, l, c, end.line, end.ecol)

    # -- module ------------------------------------------------------------
    def parse_module(self):
        while self.raw().t == "sep":
            self.p += 1
        self.expect_kw("MODULE")
        name = self.expect_id().v
        while self.raw().t == "sep":
            self.p += 1
        m = Module(name)
        while True:
            t = self.raw()
            if t.t == "eof":
                break
            if t.t == "end":
                self.p += 1
                break
            if t.t == "sep":
                self.p += 1
                if self.is_kw("MODULE"):
                    self.p -= 1
                    sub = self.parse_module()
                    m.submodules[sub.name] = sub
                continue
            self.parse_unit(m)
        return m

    def parse_unit(self, m: Module, local=False):
        t = self.raw()
        if t.t == "kw":
            kw = t.v
            if kw == "EXTENDS":
                self.p += 1
                m.extends.append(self.expect_id().v)
                while self.is_op(","):
                    self.p += 1
                    m.extends.append(self.expect_id().v)
                return
            if kw in ("CONSTANT", "CONSTANTS"):
                self.p += 1
                while True:
                    m.constants.append(self.parse_opdecl())
                    if self.is_op(","):
                        self.p += 1
                        continue
                    break
                return
            if kw in ("VARIABLE", "VARIABLES"):
                self.p += 1
                m.variables.append(self.expect_id().v)
                while self.is_op(","):
                    self.p += 1
                    m.variables.append(self.expect_id().v)
                return
            if kw in ("ASSUME", "ASSUMPTION", "AXIOM"):
                self.p += 1
                name = None
                if self.raw().t == "id" and self.raw(1).t == "op" and self.raw(1).v == "==":
                    name = self.next().v
                    self.p += 1
                e = self.parse_expr(0)
                m.assumes.append((name, e))
                if name:
                    d = OpDef(name, [], e, module=m.name)
                    m.defs[name] = d
                return
            if kw in ("THEOREM", "LEMMA", "PROPOSITION", "COROLLARY"):
                self.p += 1
                name = None
                if self.raw().t == "id" and self.raw(1).t == "op" and self.raw(1).v == "==":
                    name = self.next().v
                    self.p += 1
                e = self.parse_expr(0)
                m.theorems.append((name, e))
                if name:
                    m.defs[name] = OpDef(name, [], e, module=m.name)
                self.skip_proof()
                return
            if kw == "LOCAL":
                self.p += 1
                return self.parse_unit(m, local=True)
            if kw == "INSTANCE":
                inst = self.parse_instance(None, [])
                inst.local = local
                m.instances.append(inst)
                return
            if kw == "RECURSIVE":
                self.p += 1
                while True:
                    self.parse_opdecl()
                    if self.is_op(","):
                        self.p += 1
                        continue
                    break
                return
            raise self.err("unexpected keyword at module level")
        d = self.parse_definition(m.name)
        if isinstance(d, Instance):
            d.local = local
            m.instances.append(d)
        else:
            d.local = local
            m.defs[d.name] = d

    def skip_proof(self):
        t = self.raw()
        if not (t.t == "step" or (t.t == "kw" and t.v in ("BY", "PROOF", "OBVIOUS", "OMITTED"))):
            return
        while True:
            t = self.raw()
            if t.t in ("sep", "end", "eof"):
                return
            if t.t == "kw" and t.v in UNIT_KW and t.col == 1:
                return
            self.p += 1

    def parse_opdecl(self):
        # Name | Name(_,_) | _ op _ forms
        t = self.next()
        if t.t == "id":
            ar = 0
            if self.is_op("("):
                self.p += 1
                while True:
                    self.expect_op("_")
                    ar += 1
                    if self.is_op(","):
                        self.p += 1
                        continue
                    break
                self.expect_op(")")
            return (t.v, ar)
        if t.t == "op" and t.v == "_":
            op = self.next()
            if self.is_op("_"):
                self.p += 1
                return (op.v, 2)
            return (op.v, 1)
        if t.t == "op":  # prefix op decl  -. _
            self.expect_op("_")
            return (t.v, 1)
        raise self.err("bad operator declaration")

    def parse_instance(self, name, params):
        self.expect_kw("INSTANCE")
        mod = self.expect_id().v
        substs = []
        if self.is_kw("WITH"):
            self.p += 1
            while True:
                t = self.next()
                lhs = t.v
                self.expect_op("<-")
                substs.append((lhs, self.parse_expr(0)))
                if self.is_op(","):
                    self.p += 1
                    continue
                break
        return Instance(name, params, mod, substs)

    def parse_definition(self, modname):
        """Op == e | Op(p, q(_)) == e | f[x \\in S] == e | a (+) b == e | Name == INSTANCE ..."""
        t0 = self.raw()
        t = self.next()
        if t.t == "id":
            name = t.v
            params = []
            if self.is_op("("):
                self.p += 1
                while True:
                    params.append(self.parse_opdecl())
                    if self.is_op(","):
                        self.p += 1
                        continue
                    break
                self.expect_op(")")
                self.expect_op("==")
            elif self.is_op("["):
                # function definition
                self.p += 1
                bounds = self.parse_bounds()
                self.expect_op("]")
                self.expect_op("==")
                body = self.parse_expr(0)
                fn = self.mk("fcndef", (name, bounds, body), t0)
                d = OpDef(name, [], fn, module=modname)
                self._setloc(d, t0)
                return d
            elif self.is_op("=="):
                self.p += 1
            else:
                # infix operator definition:  a \oplus b == e
                op = self.next()
                if op.t != "op":
                    raise self.err("expected ==")
                if self.is_op("=="):
                    # postfix op def: a^+ == e
                    self.p += 1
                    body = self.parse_expr(0)
                    d = OpDef(op.v, [(name, 0)], body, module=modname)
                    self._setloc(d, t0)
                    return d
                rhs = self.expect_id().v
                self.expect_op("==")
                body = self.parse_expr(0)
                d = OpDef(op.v, [(name, 0), (rhs, 0)], body, module=modname)
                self._setloc(d, t0)
                return d
            if self.is_kw("INSTANCE"):
                return self.parse_instance(name, params)
            body = self.parse_expr(0)
            d = OpDef(name, params, body, module=modname)
            self._setloc(d, t0)
            return d
        if t.t == "op":
            # prefix operator definition  -. a == e
            if t.v == "-" and self.is_op("."):
                self.p += 1
            a = self.expect_id().v
            self.expect_op("==")
            body = self.parse_expr(0)
            d = OpDef(t.v, [(a, 0)], body, module=modname)
            self._setloc(d, t0)
            return d
        self.p -= 1
        raise self.err("expected definition")

    def _setloc(self, d, t0):
        e = self.toks[self.p - 1]
        d.line, d.col, d.eline, d.ecol = t0.line, t0.col, e.line, e.ecol

    # -- bounds ------------------------------------------------------------
    def parse_bounds(self, allow_unbounded=False):
        """x \\in S, y, z \\in T, <<a,b>> \\in U  -> [(pat, S)], pat = name | ('tuple',[names])"""
        out = []
        while True:
            pats = []
            if self.is_op("<<"):
                self.p += 1
                names = [self.expect_id().v]
                while self.is_op(","):
                    self.p += 1
                    names.append(self.expect_id().v)
                self.expect_op(">>")
                pats.append(("tuple", names))
            else:
                pats.append(self.expect_id().v)
                while self.is_op(",") and self.peek(1).t == "id" and not (
                        self.peek(2).t == "op" and self.peek(2).v in ("|->",)):
                    # x, y \in S   (but stop if this is a new bound group is handled below)
                    # lookahead: id followed by ',' or '\in'
                    nxt = self.peek(2)
                    if nxt.t == "op" and nxt.v in (",", "\\in"):
                        self.p += 1
                        pats.append(self.expect_id().v)
                    else:
                        break
            if self.is_op("\\in"):
                self.p += 1
                s = self.parse_expr(6)  # above relational so ':' / ',' end it; `..` etc absorbed
                for pt in pats:
                    out.append((pt, s))
            else:
                if not allow_unbounded:
                    raise self.err("expected \\in in bound")
                for pt in pats:
                    out.append((pt, None))
            if self.is_op(","):
                self.p += 1
                continue
            break
        return out

    def try_bounds(self):
        save = self.p
        try:
            b = self.parse_bounds()
            return b
        except ParseError:
            self.p = save
            return None

    # -- expressions -------------------------------------------------------
    def parse_expr(self, minp):
        left = self.parse_prefix()
        while True:
            t = self.peek()
            if t is not _CUT and t.t == "id":
                # instance-qualified infix operator:  a R!+ b  (Standard/Naturals.tla:6-14)
                t1 = self.toks[self.p + 1] if self.p + 1 < len(self.toks) else None
                t2 = self.toks[self.p + 2] if self.p + 2 < len(self.toks) else None
                if t1 is not None and t2 is not None and t1.t == "op" and t1.v == "!" and t2.t == "op" \
                        and t2.v in INFIX and t2.v not in ("/\\", "\\/") and INFIX[t2.v][0] >= minp:
                    lo, hi, lassoc = INFIX[t2.v]
                    self.p += 3
                    rhs = self.parse_expr(lo + 1 if lassoc else hi + 1)
                    left = self.mk("bin", (t.v + "!" + t2.v, left, rhs), left)
                    continue
            if t is _CUT or t.t != "op":
                break
            v = t.v
            if v in POSTFIX and 15 >= minp:
                self.p += 1
                left = self.mk("app", (v, (left,)), left)
                continue
            info = INFIX.get(v)
            if info is None:
                break
            lo, hi, lassoc = info
            if lo < minp:
                break
            self.p += 1
            if v == "\\X":
                items = [left, self.parse_expr(hi + 1)]
                while self.is_op("\\X"):
                    self.p += 1
                    items.append(self.parse_expr(hi + 1))
                left = self.mk("times", (tuple(items),), left)
                continue
            rhs = self.parse_expr(lo + 1 if lassoc else hi + 1)
            if v == "/\\":
                left = self.mk("and", ((left, rhs),), left)
            elif v == "\\/":
                left = self.mk("or", ((left, rhs),), left)
            else:
                if v == "=<":
                    v = "<="
                if v == "/=":
                    v = "#"
                left = self.mk("bin", (v, left, rhs), left)
        return left

    def parse_junction(self, bullet):
        t0 = self.raw()
        col = t0.col
        items = []
        while True:
            t = self.raw()
            # bullet must be visible through enclosing cuts
            if not (t.t == "op" and t.v == bullet and t.col == col):
                break
            if self.jstack and t.col <= self.jstack[-1]:
                break
            self.p += 1
            self.jstack.append(col)
            try:
                items.append(self.parse_expr(0))
            finally:
                self.jstack.pop()
        k = "and" if bullet == "/\\" else "or"
        if len(items) == 1:
            # keep single-item junction as-is (still an 'and' so Inv!1 works)
            pass
        return self.mk(k, (tuple(items),), t0)

    def parse_prefix(self):
        t = self.peek()
        if t is _CUT:
            raise self.err("expression expected")
        if t.t == "op":
            v = t.v
            if v in ("/\\", "\\/"):
                return self.parse_postfix(self.parse_junction(v))
            if v == "~":
                self.p += 1
                e = self.parse_expr(5)
                return self.mk("not", (e,), t)
            if v == "-":
                self.p += 1
                e = self.parse_expr(13)
                if e.k == "num":
                    return self.mk("num", (-e.a[0],), t)
                return self.mk("neg", (e,), t)
            if v == "[]":
                self.p += 1
                e = self.parse_expr(5)
                return self.mk("box", (e,), t)
            if v == "<>":
                self.p += 1
                e = self.parse_expr(5)
                return self.mk("diamond", (e,), t)
            if v in ("\\A", "\\E", "\\AA", "\\EE"):
                self.p += 1
                bounds = self.parse_bounds(allow_unbounded=True)
                self.expect_op(":")
                body = self.parse_expr(0)
                k = {"\\A": "forall", "\\E": "exists", "\\AA": "tforall", "\\EE": "texists"}[v]
                return self.mk(k, (bounds, body), t)
            if v in ("WF_", "SF_"):
                self.p += 1
                st = self.peek()
                if st.t == "id":
                    self.p += 1
                    sub = self.mk("id", (st.v,), st)
                else:
                    sub = self.parse_atom()
                self.expect_op("(")
                a = self.parse_expr(0)
                self.expect_op(")")
                return self.mk("wf" if v == "WF_" else "sf", (sub, a), t)
        if t.t == "kw":
            v = t.v
            if v == "IF":
                self.p += 1
                c = self.parse_expr(0)
                self.expect_kw("THEN")
                a = self.parse_expr(0)
                self.expect_kw("ELSE")
                b = self.parse_expr(0)
                return self.mk("if", (c, a, b), t)
            if v == "CASE":
                self.p += 1
                arms = []
                other = None
                while True:
                    if self.is_kw("OTHER"):
                        self.p += 1
                        self.expect_op("->")
                        other = self.parse_expr(0)
                    else:
                        c = self.parse_expr(0)
                        self.expect_op("->")
                        e = self.parse_expr(0)
                        arms.append((c, e))
                    if self.is_op("[]"):
                        self.p += 1
                        continue
                    break
                return self.mk("case", (tuple(arms), other), t)
            if v == "LET":
                self.p += 1
                defs = []
                while not self.is_kw("IN"):
                    if self.is_kw("RECURSIVE"):
                        self.p += 1
                        while True:
                            self.parse_opdecl()
                            if self.is_op(","):
                                self.p += 1
                                continue
                            break
                        continue
                    d = self.parse_definition(None)
                    defs.append(d)
                self.expect_kw("IN")
                body = self.parse_expr(0)
                return self.mk("let", (tuple(defs), body), t)
            if v == "CHOOSE":
                self.p += 1
                if self.is_op("<<"):
                    self.p += 1
                    names = [self.expect_id().v]
                    while self.is_op(","):
                        self.p += 1
                        names.append(self.expect_id().v)
                    self.expect_op(">>")
                    pat = ("tuple", names)
                else:
                    pat = self.expect_id().v
                s = None
                if self.is_op("\\in"):
                    self.p += 1
                    s = self.parse_expr(6)
                self.expect_op(":")
                body = self.parse_expr(0)
                return self.mk("choose", (pat, s, body), t)
            if v == "LAMBDA":
                self.p += 1
                params = [self.expect_id().v]
                while self.is_op(","):
                    self.p += 1
                    params.append(self.expect_id().v)
                self.expect_op(":")
                body = self.parse_expr(0)
                return self.mk("lambda", (tuple(params), body), t)
            if v in ("ENABLED", "UNCHANGED"):
                self.p += 1
                e = self.parse_expr(5)
                return self.mk(v.lower(), (e,), t)
            if v == "SUBSET":
                self.p += 1
                e = self.parse_expr(9)
                return self.mk("subset", (e,), t)
            if v == "UNION":
                self.p += 1
                e = self.parse_expr(9)
                return self.mk("bigunion", (e,), t)
            if v == "DOMAIN":
                self.p += 1
                e = self.parse_expr(10)
                return self.mk("domain", (e,), t)
        return self.parse_atom_postfix()

    def parse_atom_postfix(self):
        return self.parse_postfix(self.parse_atom())

    def parse_postfix(self, e):
        while True:
            t = self.peek()
            if t is _CUT or t.t != "op":
                break
            if t.v == "[":
                # function application (must be adjacent in intent; TLA+ has no ambiguity here
                # except `[A]_v` which never follows an expression)
                self.p += 1
                args = [self.parse_expr(0)]
                while self.is_op(","):
                    self.p += 1
                    args.append(self.parse_expr(0))
                self.expect_op("]")
                e = self.mk("fapp", (e, tuple(args)), e)
                continue
            if t.v == ".":
                self.p += 1
                f = self.next()
                if f.t not in ("id", "kw"):
                    raise self.err("field name expected")
                e = self.mk("dot", (e, f.v), e)
                continue
            if t.v == "'":
                self.p += 1
                e = self.mk("prime", (e,), e)
                continue
            break
        return e

    def parse_args(self):
        """( e1, e2 ... ) -- operator arguments; bare operator symbols allowed."""
        self.expect_op("(")
        args = []
        while True:
            t = self.peek()
            nx = self.peek(1)
            if t.t == "op" and t.v in INFIX and nx.t == "op" and nx.v in (",", ")"):
                self.p += 1
                args.append(self.mk("id", (t.v,), t))
            else:
                args.append(self.parse_expr(0))
            if self.is_op(","):
                self.p += 1
                continue
            break
        self.expect_op(")")
        return tuple(args)

    def parse_atom(self):
        t = self.next()
        if t.t == "num":
            return self.mk("num", (t.v,), t)
        if t.t == "str":
            return self.mk("str", (t.v,), t)
        if t.t == "kw":
            if t.v == "TRUE":
                return self.mk("bool", (True,), t)
            if t.v == "FALSE":
                return self.mk("bool", (False,), t)
            if t.v == "BOOLEAN":
                return self.mk("id", ("BOOLEAN",), t)
            if t.v == "STRING":
                return self.mk("id", ("STRING",), t)
            self.p -= 1
            raise self.err("unexpected keyword in expression")
        if t.t == "id":
            return self.parse_ident_path(t)
        if t.t == "op":
            v = t.v
            if v == "(":
                e = self.parse_expr(0)
                self.expect_op(")")
                return e
            if v == "<<":
                items = []
                if not (self.is_op(">>") or self.is_op(">>_")):
                    items.append(self.parse_expr(0))
                    while self.is_op(","):
                        self.p += 1
                        items.append(self.parse_expr(0))
                if self.is_op(">>_"):
                    self.p += 1
                    sub = self.parse_atom_postfix()
                    a = items[0] if len(items) == 1 else self.mk("tuple", (tuple(items),), t)
                    return self.mk("aangle", (a, sub), t)
                self.expect_op(">>")
                return self.mk("tuple", (tuple(items),), t)
            if v == "{":
                return self.parse_set(t)
            if v == "[":
                return self.parse_bracket(t)
            if v == "@":
                return self.mk("at", (), t)
        self.p -= 1
        raise self.err("unexpected token in expression")

    def parse_ident_path(self, t):
        """Ident [ (args) ] { ! (Ident[(args)] | num | : | << | >> | @) }"""
        parts = []
        name = t.v
        args = ()
        if self.is_op("("):
            args = self.parse_args()
        parts.append((name, args))
        while self.is_op("!"):
            # don't confuse with EXCEPT's `!` (only occurs after ',' or EXCEPT, never after ident)
            self.p += 1
            n = self.next()
            if n.t == "id":
                a = ()
                if self.is_op("("):
                    a = self.parse_args()
                parts.append((n.v, a))
            elif n.t == "num":
                parts.append((n.v, ()))
            elif n.t == "op" and n.v in (":", "<<", ">>", "@"):
                parts.append((n.v, ()))
            else:
                raise self.err("bad selector after !")
        if len(parts) == 1:
            if args:
                return self.mk("app", (name, args), t)
            return self.mk("id", (name,), t)
        return self.mk("sel", (tuple(parts),), t)

    def parse_set(self, t0):
        if self.is_op("}"):
            self.p += 1
            return self.mk("setenum", ((),), t0)
        # try  { x \in S : p }  /  { <<a,b>> \in S : p }
        save = self.p
        first = None
        if self.peek().t == "id" and self.is_op("\\in", 1):
            name = self.next().v
            self.p += 1
            try:
                s = self.parse_expr(6)
                if self.is_op(":"):
                    self.p += 1
                    pred = self.parse_expr(0)
                    self.expect_op("}")
                    return self.mk("setfilter", ((name, s), pred), t0)
            except ParseError:
                pass
            self.p = save
        elif self.is_op("<<"):
            b = self.try_bounds()
            if b is not None and len(b) == 1 and self.is_op(":"):
                self.p += 1
                pred = self.parse_expr(0)
                self.expect_op("}")
                return self.mk("setfilter", (b[0], pred), t0)
            self.p = save
        first = self.parse_expr(0)
        if self.is_op(":"):
            self.p += 1
            bounds = self.parse_bounds()
            self.expect_op("}")
            return self.mk("setmap", (first, bounds), t0)
        items = [first]
        while self.is_op(","):
            self.p += 1
            items.append(self.parse_expr(0))
        self.expect_op("}")
        return self.mk("setenum", (tuple(items),), t0)

    def parse_bracket(self, t0):
        # record / record set
        if self.peek().t in ("id", "kw") and self.peek(1).t == "op" and self.peek(1).v in ("|->", ":"):
            kind = self.peek(1).v
            # `[x \in S |-> e]` starts with id \in, so only id followed directly by |-> or : is a record
            pairs = []
            while True:
                f = self.next().v
                self.expect_op(kind)
                pairs.append((f, self.parse_expr(0)))
                if self.is_op(","):
                    self.p += 1
                    continue
                break
            self.expect_op("]")
            return self.mk("record" if kind == "|->" else "recset", (tuple(pairs),), t0)
        # function constructor
        save = self.p
        b = self.try_bounds()
        if b is not None and self.is_op("|->"):
            self.p += 1
            body = self.parse_expr(0)
            self.expect_op("]")
            return self.mk("fcn", (b, body), t0)
        self.p = save
        e = self.parse_expr(0)
        if self.is_op("->"):
            self.p += 1
            r = self.parse_expr(0)
            self.expect_op("]")
            return self.mk("funcset", (e, r), t0)
        if self.is_kw("EXCEPT"):
            self.p += 1
            ups = []
            while True:
                self.expect_op("!")
                path = []
                while True:
                    if self.is_op("["):
                        self.p += 1
                        idx = [self.parse_expr(0)]
                        while self.is_op(","):
                            self.p += 1
                            idx.append(self.parse_expr(0))
                        self.expect_op("]")
                        path.append(("idx", tuple(idx)))
                    elif self.is_op("."):
                        self.p += 1
                        path.append(("fld", self.next().v))
                    else:
                        break
                self.expect_op("=")
                val = self.parse_expr(0)
                ups.append((tuple(path), val))
                if self.is_op(","):
                    self.p += 1
                    continue
                break
            self.expect_op("]")
            return self.mk("except", (e, tuple(ups)), t0)
        if self.is_op("]_"):
            self.p += 1
            sub = self.parse_atom_postfix()
            return self.mk("abox", (e, sub), t0)
        raise self.err("bad bracket expression")


class _Cut:
    t = "cut"
    v = None
    line = 0
    col = 0

    def __repr__(self):
        return "CUT"


_CUT = _Cut()


def parse_module_text(text: str) -> Module:
    toks = lex(text)
    p = Parser(toks)
    return p.parse_module()


def parse_expr_text(text: str) -> Node:
    toks = lex(text, whole_file=False)
    p = Parser(toks)
    e = p.parse_expr(0)
    if p.raw().t != "eof":
        raise p.err("trailing tokens")
    return e


def read_text(path: str) -> str:
    with open(path, "rb") as f:
        data = f.read()
    try:
        return data.decode("utf-8")
    except UnicodeDecodeError:
        return data.decode("latin-1")
