"""Host-side TLA+ evaluator: constant expressions, ASSUMEs, Init enumeration and a
generic predicate solver (TLC's "x = e / x \\in S binds, everything else filters"
evaluation strategy, p-manual.pdf §4 and Paxos/MCPaxos.tla:30-35).

The product uses this module for constant evaluation, type-set evaluation and
initial-state enumeration (HOT LOOP 0 of SURVEY.md §3.2 stays on the host and is
handed to the engine through tlag_seed).  The BFS over Next is *not* done here in
the product: Next/invariants are lowered to bytecode (compile/) and run by the
CUDA engine.  oracle/tlc_oracle.py drives the same solver over Next as the
independent CPU restatement.
"""
from __future__ import annotations

import sys

from .parser import Node, OpDef
from .values import (EvalError, ModelValue, Fcn, LazyFcn, LazySet, SetNat, SetInt, SetString, SetSeq,
                     SetSubset, SetFuncs, SetRecs, SetTimes, SetUnionLazy, SetPFuncs, SetBSeq, mk_fcn, fcn_apply, fcn_domain,
                     fcn_items, is_set, is_enumerable, set_contains, set_iter, to_finite, sorted_vals,
                     values_equal, fmt, vkey)

sys.setrecursionlimit(20000)


class AssertFailure(Exception):
    def __init__(self, msg, node=None):
        super().__init__(msg)
        self.msg = msg
        self.node = node


class ModCtx:
    """A module instance: definitions + constant bindings + substitutions."""

    def __init__(self, name):
        self.name = name
        self.defs = {}       # name -> (OpDef, ModCtx)
        self.consts = {}     # name -> value
        self.const_decls = {}
        self.vars = []
        self.varset = set()
        self.substs = {}     # name -> (Node, ModCtx)
        self.instances = {}  # name -> ModCtx
        self.cache = {}
        self.assumes = []    # (name, Node)
        self.extended = set()
        self.all_instances = []  # every instantiated ctx (for <-[Mod] overrides)


class Fr:
    __slots__ = ("ctx", "s", "t")

    def __init__(self, ctx, s=None, t=None):
        self.ctx = ctx
        self.s = s
        self.t = t


class Closure:
    __slots__ = ("params", "body", "env", "ctx", "name")

    def __init__(self, params, body, env, ctx, name="LAMBDA"):
        self.params = params  # list of names
        self.body = body
        self.env = env
        self.ctx = ctx
        self.name = name


class OpVal:
    __slots__ = ("d", "ctx")

    def __init__(self, d, ctx):
        self.d = d
        self.ctx = ctx


class Thunk:
    __slots__ = ("body", "env", "ctx", "val", "done", "sdep")

    def __init__(self, body, env, ctx):
        self.body = body
        self.env = env
        self.ctx = ctx
        self.done = False
        self.val = None
        self.sdep = None


class BuiltinOp:
    __slots__ = ("name", "fn", "arity")

    def __init__(self, name, fn, arity):
        self.name = name
        self.fn = fn
        self.arity = arity


_MISSING = object()


class Evaluator:
    def __init__(self):
        self.state_reads = 0
        self.print_out = []
        self.out = sys.stdout
        self.builtins = self._make_builtins()
        self.disp = {
            "num": lambda n, e, f: n.a[0], "str": lambda n, e, f: n.a[0], "bool": lambda n, e, f: n.a[0],
            "id": self.e_id, "app": self.e_app, "sel": self.e_sel, "prime": self.e_prime,
            "tuple": self.e_tuple, "setenum": self.e_setenum, "setfilter": self.e_setfilter,
            "setmap": self.e_setmap, "fcn": self.e_fcn, "record": self.e_record, "recset": self.e_recset,
            "funcset": self.e_funcset, "except": self.e_except, "fapp": self.e_fapp, "dot": self.e_dot,
            "if": self.e_if, "case": self.e_case, "let": self.e_let, "forall": self.e_forall,
            "exists": self.e_exists, "choose": self.e_choose, "lambda": self.e_lambda,
            "and": self.e_and, "or": self.e_or, "not": self.e_not, "bin": self.e_bin, "neg": self.e_neg,
            "unchanged": self.e_unchanged, "subset": self.e_subset, "bigunion": self.e_bigunion,
            "domain": self.e_domain, "times": self.e_times, "at": self.e_at, "fcndef": self.e_fcndef,
            "abox": self.e_abox, "aangle": self.e_aangle, "enabled": self.e_unsupported,
            "box": self.e_unsupported, "diamond": self.e_unsupported, "wf": self.e_unsupported,
            "sf": self.e_unsupported, "tforall": self.e_unsupported, "texists": self.e_unsupported,
        }

    # ------------------------------------------------------------------ core
    def eval(self, n: Node, env: dict, fr: Fr):
        return self.disp[n.k](n, env, fr)

    def e_unsupported(self, n, env, fr):
        raise EvalError(f"cannot evaluate temporal/unsupported construct `{n.k}` at line {n.line}")

    def truth(self, n, env, fr):
        v = self.disp[n.k](n, env, fr)
        if v is True or v is False:
            return v
        raise EvalError(f"expected BOOLEAN at line {n.line} col {n.col}, got {fmt(v)}")

    # -- identifiers -------------------------------------------------------
    def lookup(self, name, env, fr, n=None):
        v = env.get(name, _MISSING)
        if v is not _MISSING:
            if type(v) is Thunk:
                return self.force(v, fr)
            return v
        ctx = fr.ctx
        sb = ctx.substs.get(name)
        if sb is not None:
            node, octx = sb
            return self.eval(node, {}, Fr(octx, fr.s, fr.t))
        if name in ctx.varset:
            self.state_reads += 1
            s = fr.s
            if s is None:
                raise EvalError(f"variable {name} read in a constant context")
            try:
                return s[name]
            except KeyError:
                raise EvalError(f"variable {name} is not (yet) assigned a value")
        v = ctx.consts.get(name, _MISSING)
        if v is not _MISSING:
            return v
        d = ctx.defs.get(name)
        if d is not None:
            od, dctx = d
            if od.params:
                return OpVal(od, dctx)
            return self.eval_def0(name, od, dctx, fr)
        b = self.builtins.get(name)
        if b is not None:
            if b.arity == 0:
                return b.fn()
            return b
        where = f" at line {n.line} col {n.col}" if n is not None else ""
        raise EvalError(f"unknown identifier {name}{where} (module {ctx.name})")

    def eval_def0(self, name, od, dctx, fr):
        c = dctx.cache
        v = c.get(name, _MISSING)
        if v is not _MISSING:
            return v
        before = self.state_reads
        v = self.eval(od.body, {}, Fr(dctx, fr.s, fr.t))
        if self.state_reads == before and not isinstance(v, (OpVal, Closure)):
            c[name] = v
        return v

    def force(self, th: Thunk, fr):
        if th.done and (th.sdep is None or (th.sdep[0] is fr.s and th.sdep[1] is fr.t)):
            return th.val
        before = self.state_reads
        v = self.eval(th.body, th.env, Fr(th.ctx, fr.s, fr.t))
        th.val = v
        th.done = True
        th.sdep = None if self.state_reads == before else (fr.s, fr.t)
        return v

    def e_id(self, n, env, fr):
        return self.lookup(n.a[0], env, fr, n)

    def e_at(self, n, env, fr):
        v = env.get("@", _MISSING)
        if v is _MISSING:
            raise EvalError("@ used outside EXCEPT")
        if type(v) is Thunk:
            return self.force(v, fr)
        return v

    def resolve_callable(self, name, env, fr, n=None):
        v = env.get(name, _MISSING)
        if v is not _MISSING:
            if type(v) is Thunk:
                v = self.force(v, fr)
            return v
        ctx = fr.ctx
        d = ctx.defs.get(name)
        if d is not None:
            return OpVal(d[0], d[1])
        b = self.builtins.get(name)
        if b is not None:
            return b
        sb = ctx.substs.get(name)
        if sb is not None:
            node, octx = sb
            return self.eval(node, {}, Fr(octx, fr.s, fr.t))
        c = ctx.consts.get(name, _MISSING)
        if c is not _MISSING:
            return c
        raise EvalError(f"unknown operator {name}" + (f" at line {n.line}" if n else ""))

    def eval_arg(self, a, env, fr):
        return self.disp[a.k](a, env, fr)

    def e_app(self, n, env, fr):
        name, args = n.a
        op = self.resolve_callable(name, env, fr, n)
        if type(op) is BuiltinOp:
            if name == "Assert" or name == "Print" or name == "PrintT":
                return op.fn(*[self.eval_arg(a, env, fr) for a in args], node=n)
            return op.fn(*[self.eval_arg(a, env, fr) for a in args])
        if type(op) is OpVal and fr.t is not None and len(op.d.params) == len(args):
            # inside an action, parameters are passed by name (a callee may prime them: AlternatingBit.tla Lose(q))
            argv = [Thunk(a, env, fr.ctx) if (pa[1] == 0 and a.k in ("id", "prime", "fapp", "dot")) else
                    self.eval_arg(a, env, fr) for a, pa in zip(args, op.d.params)]
            return self.apply_op(op, argv, fr, n)
        return self.apply_op(op, [self.eval_arg(a, env, fr) for a in args], fr, n)

    def apply_op(self, op, args, fr, n=None):
        t = type(op)
        if t is OpVal:
            d = op.d
            if len(d.params) != len(args):
                raise EvalError(f"operator {d.name} expects {len(d.params)} args, got {len(args)}")
            env2 = {}
            for (pn, _), a in zip(d.params, args):
                env2[pn] = a
            return self.eval(d.body, env2, Fr(op.ctx, fr.s, fr.t))
        if t is Closure:
            if len(op.params) != len(args):
                raise EvalError(f"operator {op.name} expects {len(op.params)} args, got {len(args)}")
            env2 = dict(op.env)
            for pn, a in zip(op.params, args):
                env2[pn] = a
            return self.eval(op.body, env2, Fr(op.ctx, fr.s, fr.t))
        if t is BuiltinOp:
            return op.fn(*args)
        raise EvalError(f"applying a non-operator {fmt(op)}" + (f" at line {n.line}" if n else ""))

    # -- !-selectors ---------------------------------------------------------
    def resolve_sel(self, parts, env, fr):
        """Returns (kind, payload): ('def', OpDef, ctx, argnodes) or ('expr', Node, ctx, env)."""
        ctx = fr.ctx
        i = 0
        while i < len(parts):
            name, args = parts[i]
            if isinstance(name, str) and name in ctx.instances and i + 1 < len(parts):
                ctx = ctx.instances[name]
                i += 1
                continue
            break
        name, args = parts[i]
        d = ctx.defs.get(name)
        if d is None:
            if i == len(parts) - 1:
                # could be a variable / constant reached through an instance: V!maxBal
                return ("name", name, ctx, args)
            raise EvalError(f"unknown definition {name} in selector")
        od, dctx = d
        if i == len(parts) - 1:
            return ("def", od, dctx, args)
        # label / positional selectors into the body
        node = od.body
        penv_args = args
        for sel, _ in parts[i + 1:]:
            if sel == ":":
                continue
            if isinstance(sel, int):
                node = self._nth_operand(node, sel)
            else:
                raise EvalError(f"unsupported selector !{sel}")
        return ("expr", node, dctx, (od, penv_args))

    def _nth_operand(self, node, k):
        if node.k in ("and", "or"):
            items = node.a[0]
            return items[k - 1]
        if node.k == "bin":
            return node.a[k]
        if node.k in ("not", "neg", "box", "diamond", "subset", "domain", "bigunion", "unchanged"):
            return node.a[0]
        if node.k == "if":
            return node.a[k - 1]
        if node.k in ("forall", "exists"):
            return node.a[1]
        raise EvalError(f"cannot select operand {k} of {node.k}")

    def e_sel(self, n, env, fr):
        r = self.resolve_sel(n.a[0], env, fr)
        if r[0] == "def":
            _, od, dctx, args = r
            if od.params:
                if not args:
                    return OpVal(od, dctx)
                return self.apply_op(OpVal(od, dctx), [self.eval_arg(a, env, fr) for a in args], fr, n)
            return self.eval_def0(od.name, od, dctx, fr)
        if r[0] == "name":
            _, name, ctx, args = r
            return self.lookup(name, {}, Fr(ctx, fr.s, fr.t), n)
        _, node, dctx, (od, args) = r
        env2 = {}
        if od.params:
            for (pn, _), a in zip(od.params, args):
                env2[pn] = self.eval_arg(a, env, fr)
        return self.eval(node, env2, Fr(dctx, fr.s, fr.t))

    # -- state ---------------------------------------------------------------
    def e_prime(self, n, env, fr):
        if fr.t is None:
            raise EvalError(f"primed expression at line {n.line} evaluated outside an action")
        return self.eval(n.a[0], env, Fr(fr.ctx, fr.t, None))

    def e_unchanged(self, n, env, fr):
        e = n.a[0]
        a = self.eval(e, env, fr)
        b = self.eval(e, env, Fr(fr.ctx, fr.t, None))
        return values_equal(a, b)

    def e_abox(self, n, env, fr):
        a, v = n.a
        if self.truth(a, env, fr):
            return True
        return values_equal(self.eval(v, env, fr), self.eval(v, env, Fr(fr.ctx, fr.t, None)))

    def e_aangle(self, n, env, fr):
        a, v = n.a
        if not self.truth(a, env, fr):
            return False
        return not values_equal(self.eval(v, env, fr), self.eval(v, env, Fr(fr.ctx, fr.t, None)))

    # -- constructors ------------------------------------------------------
    def e_tuple(self, n, env, fr):
        return tuple(self.eval(x, env, fr) for x in n.a[0])

    def e_setenum(self, n, env, fr):
        return frozenset(self.eval(x, env, fr) for x in n.a[0])

    def bind_pat(self, env, pat, v):
        if isinstance(pat, str):
            env[pat] = v
        else:
            names = pat[1]
            if not isinstance(v, tuple) or len(v) != len(names):
                raise EvalError(f"tuple pattern <<{', '.join(names)}>> does not match {fmt(v)}")
            for nm, x in zip(names, v):
                env[nm] = x

    def iter_bounds(self, bounds, env, fr, i=0):
        """Yield extended envs for all bound combinations in canonical order."""
        if i == len(bounds):
            yield env
            return
        pat, sn = bounds[i]
        if sn is None:
            raise EvalError("unbounded quantifier cannot be evaluated")
        s = self.eval(sn, env, fr)
        for v in set_iter(s):
            env2 = dict(env)
            self.bind_pat(env2, pat, v)
            yield from self.iter_bounds(bounds, env2, fr, i + 1)

    def e_setfilter(self, n, env, fr):
        (pat, sn), pred = n.a
        s = self.eval(sn, env, fr)
        out = []
        env2 = dict(env)
        for v in set_iter(s):
            self.bind_pat(env2, pat, v)
            if self.truth(pred, env2, fr):
                out.append(v)
        return frozenset(out)

    def e_setmap(self, n, env, fr):
        e, bounds = n.a
        return frozenset(self.eval(e, env2, fr) for env2 in self.iter_bounds(bounds, env, fr))

    def e_fcn(self, n, env, fr):
        bounds, body = n.a
        if len(bounds) == 1:
            pat, sn = bounds[0]
            s = self.eval(sn, env, fr)
            d = {}
            env2 = dict(env)
            for v in set_iter(s):
                self.bind_pat(env2, pat, v)
                d[v] = self.eval(body, env2, fr)
            return mk_fcn(d)
        d = {}
        for env2, key in self._iter_bounds_keys(bounds, env, fr):
            d[key] = self.eval(body, env2, fr)
        return mk_fcn(d)

    def _iter_bounds_keys(self, bounds, env, fr, i=0, key=()):
        if i == len(bounds):
            yield env, key
            return
        pat, sn = bounds[i]
        s = self.eval(sn, env, fr)
        for v in set_iter(s):
            env2 = dict(env)
            self.bind_pat(env2, pat, v)
            yield from self._iter_bounds_keys(bounds, env2, fr, i + 1, key + (v,))

    def e_fcndef(self, n, env, fr):
        name, bounds, body = n.a
        ctx = fr.ctx
        s_, t_ = fr.s, fr.t
        env2 = dict(env)

        def dom():
            if len(bounds) == 1:
                return self.eval(bounds[0][1], env2, Fr(ctx, s_, t_))
            return SetTimes([self.eval(b[1], env2, Fr(ctx, s_, t_)) for b in bounds])

        def app(arg):
            e3 = dict(env2)
            if len(bounds) == 1:
                self.bind_pat(e3, bounds[0][0], arg)
            else:
                for (pat, _), x in zip(bounds, arg):
                    self.bind_pat(e3, pat, x)
            return self.eval(body, e3, Fr(ctx, s_, t_))

        lf = LazyFcn(dom, app)
        env2[name] = lf
        return lf

    def e_record(self, n, env, fr):
        return Fcn({f: self.eval(e, env, fr) for f, e in n.a[0]})

    def e_recset(self, n, env, fr):
        return SetRecs([(f, self.eval(e, env, fr)) for f, e in n.a[0]])

    def e_funcset(self, n, env, fr):
        return SetFuncs(self.eval(n.a[0], env, fr), self.eval(n.a[1], env, fr))

    def e_times(self, n, env, fr):
        return SetTimes([self.eval(x, env, fr) for x in n.a[0]])

    def e_subset(self, n, env, fr):
        return SetSubset(self.eval(n.a[0], env, fr))

    def e_bigunion(self, n, env, fr):
        lz = self._structured_union(n.a[0], env, fr)
        if lz is not None:
            return lz
        s = self.eval(n.a[0], env, fr)
        out = set()
        for x in set_iter(s):
            out |= to_finite(x)
        return frozenset(out)

    def _structured_union(self, m, env, fr):
        """Two unions that stay lazy (their element count is astronomical for typing definitions such as
        raft's voterLog / mlog):  UNION {[d -> R] : d \\in SUBSET D}  and  UNION {[1..k -> S] : k \\in 0..n}."""
        if m.k != "setmap" or len(m.a[1]) != 1:
            return None
        body, ((pat, sn),) = m.a
        if not isinstance(pat, str) or body.k != "funcset" or sn is None:
            return None
        dn, rn = body.a

        def mentions(x):
            if isinstance(x, Node):
                return (x.k == "id" and x.a[0] == pat) or any(mentions(y) for y in x.a)
            if isinstance(x, (tuple, list)):
                return any(mentions(y) for y in x)
            return False
        if mentions(rn):
            return None
        if sn.k == "subset" and dn.k == "id" and dn.a[0] == pat:
            return SetPFuncs(self.eval(sn.a[0], env, fr), self.eval(rn, env, fr))
        if sn.k == "bin" and sn.a[0] == ".." and dn.k == "bin" and dn.a[0] == ".." \
                and dn.a[2].k == "id" and dn.a[2].a[0] == pat and not mentions(dn.a[1]):
            lo = self.eval(sn.a[1], env, fr)
            one = self.eval(dn.a[1], env, fr)
            hi = self.eval(sn.a[2], env, fr)
            if lo == 0 and one == 1 and type(hi) is int:
                return SetBSeq(self.eval(rn, env, fr), hi)
        return None

    def e_domain(self, n, env, fr):
        return fcn_domain(self.eval(n.a[0], env, fr))

    def e_fapp(self, n, env, fr):
        f = self.eval(n.a[0], env, fr)
        args = n.a[1]
        if len(args) == 1:
            arg = self.eval(args[0], env, fr)
        else:
            arg = tuple(self.eval(a, env, fr) for a in args)
        try:
            return fcn_apply(f, arg)
        except EvalError as ex:
            raise EvalError(f"{ex} (line {n.line} col {n.col})")

    def e_dot(self, n, env, fr):
        r = self.eval(n.a[0], env, fr)
        if isinstance(r, Fcn):
            try:
                return r.d[n.a[1]]
            except KeyError:
                pass
        raise EvalError(f"record {fmt(r)} has no field {n.a[1]} (line {n.line} col {n.col})")

    def e_except(self, n, env, fr):
        f = self.eval(n.a[0], env, fr)
        for path, val in n.a[1]:
            f = self._except(f, path, 0, val, env, fr)
        return f

    def _except(self, f, path, i, valnode, env, fr):
        if i == len(path):
            env2 = dict(env)
            env2["@"] = f
            return self.eval(valnode, env2, fr)
        kind, p = path[i]
        if isinstance(f, LazyFcn):
            f = f.force()
        if kind == "fld":
            key = p
        else:
            key = self.eval(p[0], env, fr) if len(p) == 1 else tuple(self.eval(x, env, fr) for x in p)
        if isinstance(f, tuple):
            if not (type(key) is int and 1 <= key <= len(f)):
                return f  # TLC: EXCEPT outside domain leaves f unchanged (with a warning)
            new = self._except(f[key - 1], path, i + 1, valnode, env, fr)
            return f[:key - 1] + (new,) + f[key:]
        if isinstance(f, Fcn):
            if key not in f.d:
                return f
            d = dict(f.d)
            d[key] = self._except(f.d[key], path, i + 1, valnode, env, fr)
            return Fcn(d)
        raise EvalError(f"EXCEPT applied to non-function {fmt(f)}")

    # -- control -------------------------------------------------------------
    def e_if(self, n, env, fr):
        if self.truth(n.a[0], env, fr):
            return self.eval(n.a[1], env, fr)
        return self.eval(n.a[2], env, fr)

    def e_case(self, n, env, fr):
        arms, other = n.a
        for c, e in arms:
            if self.truth(c, env, fr):
                return self.eval(e, env, fr)
        if other is not None:
            return self.eval(other, env, fr)
        raise EvalError(f"CASE at line {n.line}: no arm is true")

    def let_env(self, defs, env, fr):
        env2 = dict(env)
        for d in defs:
            if d.params:
                env2[d.name] = Closure([p for p, _ in d.params], d.body, env2, fr.ctx, d.name)
            else:
                env2[d.name] = Thunk(d.body, env2, fr.ctx)
        return env2

    def e_let(self, n, env, fr):
        defs, body = n.a
        return self.eval(body, self.let_env(defs, env, fr), fr)

    def e_lambda(self, n, env, fr):
        return Closure(list(n.a[0]), n.a[1], env, fr.ctx)

    def e_forall(self, n, env, fr):
        bounds, body = n.a
        for env2 in self.iter_bounds(bounds, env, fr):
            if not self.truth(body, env2, fr):
                return False
        return True

    def e_exists(self, n, env, fr):
        bounds, body = n.a
        for env2 in self.iter_bounds(bounds, env, fr):
            if self.truth(body, env2, fr):
                return True
        return False

    def e_choose(self, n, env, fr):
        pat, sn, body = n.a
        if sn is None:
            raise EvalError(f"unbounded CHOOSE at line {n.line} cannot be evaluated "
                            f"(bind the defined symbol to a model value in the cfg)")
        s = self.eval(sn, env, fr)
        env2 = dict(env)
        for v in set_iter(s):
            self.bind_pat(env2, pat, v)
            if self.truth(body, env2, fr):
                return v
        raise EvalError(f"CHOOSE at line {n.line}: no element of {fmt(s)} satisfies the predicate")

    def e_and(self, n, env, fr):
        for x in n.a[0]:
            if not self.truth(x, env, fr):
                return False
        return True

    def e_or(self, n, env, fr):
        for x in n.a[0]:
            if self.truth(x, env, fr):
                return True
        return False

    def e_not(self, n, env, fr):
        return not self.truth(n.a[0], env, fr)

    def e_neg(self, n, env, fr):
        v = self.eval(n.a[0], env, fr)
        if type(v) is not int:
            raise EvalError(f"unary minus on non-integer {fmt(v)}")
        return -v

    def e_bin(self, n, env, fr):
        op, ln, rn = n.a
        if op == "=>":
            return (not self.truth(ln, env, fr)) or self.truth(rn, env, fr)
        if op == "<=>":
            return self.truth(ln, env, fr) == self.truth(rn, env, fr)
        # user-defined infix operator?
        if op not in _BIN_BUILTIN or op in fr.ctx.defs or op in env:
            cal = None
            if op in env:
                cal = env[op]
            elif op in fr.ctx.defs:
                cal = OpVal(*fr.ctx.defs[op])
            if cal is not None:
                return self.apply_op(cal, [self.eval(ln, env, fr), self.eval(rn, env, fr)], fr, n)
            if op not in _BIN_BUILTIN:
                raise EvalError(f"undefined infix operator {op} at line {n.line}")
        a = self.eval(ln, env, fr)
        b = self.eval(rn, env, fr)
        return self.binop(op, a, b, n)

    def binop(self, op, a, b, n=None):
        if op == "=":
            return values_equal(a, b)
        if op == "#":
            return not values_equal(a, b)
        if op == "\\in":
            return set_contains(b, a)
        if op == "\\notin":
            return not set_contains(b, a)
        if op in ("+", "-", "*", "<", ">", "<=", ">=", "\\div", "%", "..", "^"):
            if type(a) is not int or type(b) is not int:
                raise EvalError(f"operator {op} applied to non-integers {fmt(a)}, {fmt(b)}"
                                + (f" at line {n.line} col {n.col}" if n else ""))
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "<":
                return a < b
            if op == ">":
                return a > b
            if op == "<=":
                return a <= b
            if op == ">=":
                return a >= b
            if op == "\\div":
                if b == 0:
                    raise EvalError("division by zero")
                return a // b
            if op == "%":
                if b <= 0:
                    raise EvalError("modulus by non-positive number")
                return a % b
            if op == "..":
                return frozenset(range(a, b + 1))
            if op == "^":
                return a ** b
        if op == "\\cup":
            if isinstance(a, frozenset) and isinstance(b, frozenset):
                return a | b
            if not (is_set(a) and is_set(b)):
                raise EvalError(f"\\cup applied to non-sets {fmt(a)}, {fmt(b)}")
            return SetUnionLazy(a, b)
        if op == "\\cap":
            if isinstance(a, frozenset) and isinstance(b, frozenset):
                return a & b
            if is_enumerable(a):
                return frozenset(x for x in set_iter(a) if set_contains(b, x))
            if is_enumerable(b):
                return frozenset(x for x in set_iter(b) if set_contains(a, x))
            raise EvalError("\\cap of two non-enumerable sets")
        if op == "\\":
            if isinstance(a, frozenset) and isinstance(b, frozenset):
                return a - b
            return frozenset(x for x in set_iter(a) if not set_contains(b, x))
        if op == "\\subseteq":
            if isinstance(a, frozenset) and isinstance(b, frozenset):
                return a <= b
            return all(set_contains(b, x) for x in set_iter(a))
        if op == "\\subset":
            return all(set_contains(b, x) for x in set_iter(a)) and not values_equal(a, b)
        if op == "\\supseteq":
            return all(set_contains(a, x) for x in set_iter(b))
        if op == "\\o":
            if not (isinstance(a, tuple) and isinstance(b, tuple)):
                raise EvalError(f"\\o applied to non-sequences {fmt(a)}, {fmt(b)}")
            return a + b
        if op == ":>":
            return mk_fcn({a: b})
        if op == "@@":
            d = dict(fcn_items(b))
            d.update(dict(fcn_items(a)))
            return mk_fcn(d)
        raise EvalError(f"unsupported operator {op}")

    # ------------------------------------------------------------- builtins
    def _make_builtins(self):
        B = {}

        def reg(name, arity):
            def deco(fn):
                B[name] = BuiltinOp(name, fn, arity)
                return fn
            return deco

        reg("Nat", 0)(lambda: SetNat())
        reg("Int", 0)(lambda: SetInt())
        reg("BOOLEAN", 0)(lambda: frozenset((False, True)))
        reg("STRING", 0)(lambda: SetString())

        @reg("Len", 1)
        def _len(s):
            if not isinstance(s, tuple):
                raise EvalError(f"Len of non-sequence {fmt(s)}")
            return len(s)

        @reg("Append", 2)
        def _append(s, e):
            if not isinstance(s, tuple):
                raise EvalError(f"Append to non-sequence {fmt(s)}")
            return s + (e,)

        @reg("Head", 1)
        def _head(s):
            if not isinstance(s, tuple) or not s:
                raise EvalError("Head of empty/non sequence")
            return s[0]

        @reg("Tail", 1)
        def _tail(s):
            if not isinstance(s, tuple) or not s:
                raise EvalError("Tail of empty/non sequence")
            return s[1:]

        @reg("SubSeq", 3)
        def _subseq(s, m, n):
            if not isinstance(s, tuple):
                raise EvalError("SubSeq of non-sequence")
            if m > n:
                return ()
            if m < 1 or n > len(s):
                raise EvalError(f"SubSeq({fmt(s)}, {m}, {n}) out of range")
            return s[m - 1:n]

        @reg("SelectSeq", 2)
        def _selectseq(s, test):
            return tuple(x for x in s if self.apply_op(test, [x], self._cur_fr) is True)

        reg("Seq", 1)(lambda s: SetSeq(s))

        @reg("Cardinality", 1)
        def _card(s):
            return len(to_finite(s))

        reg("IsFiniteSet", 1)(lambda s: isinstance(s, frozenset) or (isinstance(s, LazySet) and s.finite))

        @reg("Permutations", 1)
        def _perms(s):
            import itertools
            base = sorted_vals(to_finite(s))
            return frozenset(mk_fcn(dict(zip(base, p))) for p in itertools.permutations(base))

        @reg("SortSeq", 2)
        def _sortseq(s, lt):
            import functools

            def cmp(a, b):
                if self.apply_op(lt, [a, b], self._cur_fr):
                    return -1
                if self.apply_op(lt, [b, a], self._cur_fr):
                    return 1
                return 0
            return tuple(sorted(s, key=functools.cmp_to_key(cmp)))

        def _print(out, val, node=None):
            self.print_out.append(fmt(out))
            print(fmt(out), file=self.out)
            return val
        B["Print"] = BuiltinOp("Print", _print, 2)

        def _printt(out, node=None):
            self.print_out.append(fmt(out))
            print(fmt(out), file=self.out)
            return True
        B["PrintT"] = BuiltinOp("PrintT", _printt, 1)

        def _assert(cond, msg, node=None):
            if cond is True:
                return True
            raise AssertFailure(msg, node)
        B["Assert"] = BuiltinOp("Assert", _assert, 2)

        reg("JavaTime", 0)(lambda: 0)
        reg("ToString", 1)(lambda v: fmt(v))
        reg("TLCGet", 1)(lambda i: 0)
        reg("TLCSet", 2)(lambda i, v: True)
        # Bags (Standard/Bags.tla) -- minimal
        reg("EmptyBag", 0)(lambda: ())
        reg("SetToBag", 1)(lambda s: mk_fcn({x: 1 for x in to_finite(s)}))
        reg("BagToSet", 1)(lambda b: fcn_domain(b))
        return B

    _cur_fr = Fr(None)

    # ---------------------------------------------------------------- solver
    def solve(self, n, env, ctx, s, asg, target, act=None):
        """Enumerate assignments satisfying predicate n.

        target 'next': binds primed variables (x' = e, x' \\in S, UNCHANGED) into asg,
                       current state s is read-only.
        target 'cur' : binds unprimed variables (Init) into asg.
        Yields (asg, act) with asg a (possibly shared) dict; callers must not mutate."""
        k = n.k
        if k == "and":
            yield from self._solve_seq(n.a[0], 0, env, ctx, s, asg, target, act)
            return
        if k == "or":
            for x in n.a[0]:
                yield from self.solve(x, env, ctx, s, asg, target, act)
            return
        if k == "exists":
            bounds, body = n.a
            fr = self._fr(ctx, s, asg, target)
            for env2 in self.iter_bounds(bounds, env, fr):
                yield from self.solve(body, env2, ctx, s, asg, target, act)
            return
        if k == "if":
            fr = self._fr(ctx, s, asg, target)
            if self.truth(n.a[0], env, fr):
                yield from self.solve(n.a[1], env, ctx, s, asg, target, act)
            else:
                yield from self.solve(n.a[2], env, ctx, s, asg, target, act)
            return
        if k == "case":
            fr = self._fr(ctx, s, asg, target)
            arms, other = n.a
            for c, e in arms:
                if self.truth(c, env, fr):
                    yield from self.solve(e, env, ctx, s, asg, target, act)
                    return
            if other is not None:
                yield from self.solve(other, env, ctx, s, asg, target, act)
                return
            raise EvalError(f"CASE at line {n.line}: no arm is true")
        if k == "let":
            fr = self._fr(ctx, s, asg, target)
            env2 = self.let_env(n.a[0], env, fr)
            yield from self.solve(n.a[1], env2, ctx, s, asg, target, act)
            return
        if k == "id" or k == "app":
            name = n.a[0]
            args = n.a[1] if k == "app" else ()
            cal = None
            v = env.get(name, _MISSING)
            if v is not _MISSING:
                if type(v) is Closure:
                    cal = v
                elif type(v) is Thunk:
                    # LET-defined action without params
                    yield from self.solve(v.body, v.env, v.ctx, s, asg, target, act)
                    return
                elif type(v) is OpVal:
                    cal = v
            elif name in ctx.defs and name not in ctx.varset:
                cal = OpVal(*ctx.defs[name])
            if cal is not None and not (type(cal) is OpVal and cal.d.name in ("Assert",)):
                fr = self._fr(ctx, s, asg, target)
                # call-by-name for action parameters so that `Send(p, d, memInt, memInt')`
                # (CachingMemory/MCInternalMemory.tla:15-29) can bind memInt' inside the callee
                argv = [Thunk(a, env, ctx) if a.k in ("prime", "id", "fapp", "app", "dot") else
                        self.eval_arg(a, env, fr) for a in args]
                if type(cal) is OpVal:
                    argv = [self.force(a, fr) if type(a) is Thunk and pa[1] > 0 else a
                            for a, pa in zip(argv, cal.d.params)] if len(argv) == len(cal.d.params) else argv
                    d = cal.d
                    if len(d.params) != len(argv):
                        raise EvalError(f"operator {d.name} arity mismatch at line {n.line}")
                    env2 = {pn: a for (pn, _), a in zip(d.params, argv)}
                    if act is None or act[0] == "split":
                        act2 = ("split", d.name, d.body.loc(), cal.ctx.name)
                    else:
                        act2 = act
                    yield from self.solve(d.body, env2, cal.ctx, s, asg, target, act2)
                else:
                    env2 = dict(cal.env)
                    for pn, a in zip(cal.params, argv):
                        env2[pn] = a
                    yield from self.solve(cal.body, env2, cal.ctx, s, asg, target, act)
                return
        if k == "sel":
            fr = self._fr(ctx, s, asg, target)
            r = self.resolve_sel(n.a[0], env, fr)
            if r[0] == "def":
                _, od, dctx, args = r
                env2 = {pn: self.eval_arg(a, env, fr) for (pn, _), a in zip(od.params, args)}
                yield from self.solve(od.body, env2, dctx, s, asg, target, act)
                return
            if r[0] == "expr":
                _, node, dctx, (od, args) = r
                env2 = {pn: self.eval_arg(a, env, fr) for (pn, _), a in zip(od.params, args)}
                yield from self.solve(node, env2, dctx, s, asg, target, act)
                return
        if k == "bin":
            op, ln, rn = n.a
            if op == "=" or op == "\\in":
                var = self._assign_target(ln, ctx, asg, target, env)
                if var is not None:
                    fr = self._fr(ctx, s, asg, target)
                    rv = self.eval(rn, env, fr)
                    if isinstance(rv, LazyFcn):
                        rv = rv.force()
                    if op == "=":
                        if isinstance(rv, LazySet):
                            rv = to_finite(rv)
                        a2 = dict(asg)
                        a2[var] = rv
                        yield a2, act
                    else:
                        for v in set_iter(rv):
                            a2 = dict(asg)
                            a2[var] = v
                            yield a2, act
                    return
        if k == "unchanged" and target == "next":
            yield from self._solve_unchanged(n.a[0], env, ctx, s, asg, act)
            return
        if k == "abox" and target == "next":
            # [A]_v used as an action (MCRealTimeHourClock.tla:23-24): A \/ UNCHANGED v
            yield from self.solve(n.a[0], env, ctx, s, asg, target, act)
            yield from self._solve_unchanged(n.a[1], env, ctx, s, asg, act)
            return
        # plain boolean filter
        fr = self._fr(ctx, s, asg, target)
        if self.truth(n, env, fr):
            yield asg, act

    def _fr(self, ctx, s, asg, target):
        fr = Fr(ctx, s, asg) if target == "next" else Fr(ctx, asg, None)
        self._cur_fr = fr
        return fr

    def _solve_seq(self, items, i, env, ctx, s, asg, target, act):
        if len(items) == 1:
            yield from self.solve(items[0], env, ctx, s, asg, target, act)
            return
        if act is not None and act[0] == "split":
            act = ("fixed",) + act[1:]
        yield from self._seq(items, 0, env, ctx, s, asg, target, act)

    def _seq(self, items, i, env, ctx, s, asg, target, act):
        if i == len(items):
            yield asg, act
            return
        for a2, _ in self.solve(items[i], env, ctx, s, asg, target, act):
            yield from self._seq(items, i + 1, env, ctx, s, a2, target, act)

    def _assign_target(self, ln, ctx, asg, target, env):
        if ln.k == "id":
            th = env.get(ln.a[0])
            if type(th) is Thunk and not th.done and th.body.k in ("prime", "id"):
                return self._assign_target(th.body, th.ctx, asg, target, th.env)
        if ln.k == "prime" and ln.a[0].k == "id":
            # q' where q is an operator parameter passed a variable by name: Lose(q) == ... q' = ...
            # (TLC/AlternatingBit.tla) -- the primed parameter denotes the primed argument
            th = env.get(ln.a[0].a[0])
            if type(th) is Thunk and th.body.k == "id":
                return self._assign_target(Node("prime", (th.body,)), th.ctx, asg, target, th.env)
        if target == "next":
            if ln.k == "prime" and ln.a[0].k == "id":
                v = ln.a[0].a[0]
                if v in ctx.varset and v not in env and v not in ctx.substs and v not in asg:
                    return v
            return None
        if ln.k == "id":
            v = ln.a[0]
            if v in ctx.varset and v not in env and v not in ctx.substs and v not in asg:
                return v
        return None

    def _solve_unchanged(self, e, env, ctx, s, asg, act):
        """UNCHANGED <<x, y>> / UNCHANGED vars: bind x' = x for unassigned vars."""
        names = []
        if not self._unchanged_vars(e, env, ctx, names):
            fr = Fr(ctx, s, asg)
            a = self.eval(e, env, fr)
            b = self.eval(e, env, Fr(ctx, asg, None))
            if values_equal(a, b):
                yield asg, act
            return
        a2 = None
        for v in names:
            if v in asg:
                if not values_equal(asg[v], s[v]):
                    return
            else:
                if a2 is None:
                    a2 = dict(asg)
                a2[v] = s[v]
        yield (a2 if a2 is not None else asg), act

    def _unchanged_vars(self, e, env, ctx, out):
        if e.k == "id":
            nm = e.a[0]
            if nm in env:
                return False
            if nm in ctx.varset and nm not in ctx.substs:
                out.append(nm)
                return True
            d = ctx.defs.get(nm)
            if d is not None and not d[0].params and d[1] is ctx:
                return self._unchanged_vars(d[0].body, env, ctx, out)
            return False
        if e.k == "tuple":
            return all(self._unchanged_vars(x, env, ctx, out) for x in e.a[0])
        return False


_BIN_BUILTIN = {"=", "#", "\\in", "\\notin", "+", "-", "*", "<", ">", "<=", ">=", "\\div", "%", "..", "^",
                "\\cup", "\\cap", "\\", "\\subseteq", "\\subset", "\\supseteq", "\\o", ":>", "@@"}
