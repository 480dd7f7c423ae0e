"""TLA+ value universe for the host-side evaluator.

Semantics follow the standard-module definitions the reference ships:
examples/SpecifyingSystems/Standard/Naturals.tla:4-16, Integers.tla:5-6,
Sequences.tla:14-58, FiniteSets.tla:9-22 and TLC/TLC.tla:5-24.

Representation
  integers      Python int            booleans  Python bool
  strings       Python str            model values  ModelValue (interned)
  finite sets   frozenset             tuples / sequences  Python tuple
  functions / records   Fcn (immutable mapping).  A function whose domain is
                1..n is *normalised to a tuple* so that <<a,b>> = [i \\in 1..2 |-> ..]
                and [j \\in {} |-> e] = <<>> hold (raft.tla:158,190 rely on it).
  big / infinite sets   lazy classes (Nat, Int, Seq(S), SUBSET S, [S -> T], ...)
"""
from __future__ import annotations

import itertools


class EvalError(Exception):
    pass


class ModelValue:
    __slots__ = ("name",)
    _pool: dict = {}

    def __new__(cls, name):
        mv = cls._pool.get(name)
        if mv is None:
            mv = object.__new__(cls)
            mv.name = name
            cls._pool[name] = mv
        return mv

    def __repr__(self):
        return self.name

    def __reduce__(self):
        return (ModelValue, (self.name,))


class Fcn:
    """Immutable finite function (records are functions over strings)."""
    __slots__ = ("d", "_h")

    def __init__(self, d: dict):
        self.d = d
        self._h = None

    def __hash__(self):
        h = self._h
        if h is None:
            h = self._h = hash(frozenset(self.d.items()))
        return h

    def __eq__(self, o):
        return isinstance(o, Fcn) and self.d == o.d

    def __ne__(self, o):
        return not self.__eq__(o)

    def __repr__(self):
        return fmt(self)


def mk_fcn(d: dict):
    """Build a function value, normalising 1..n-domain functions to tuples."""
    n = len(d)
    if n == 0:
        return ()
    if all(type(k) is int for k in d):
        if min(d) == 1 and max(d) == n:
            return tuple(d[i] for i in range(1, n + 1))
    return Fcn(d)


def is_fcn_like(v):
    return isinstance(v, (Fcn, tuple))


def fcn_domain(v):
    if isinstance(v, tuple):
        return frozenset(range(1, len(v) + 1))
    if isinstance(v, Fcn):
        return frozenset(v.d.keys())
    if isinstance(v, LazyFcn):
        return v.domain()
    raise EvalError(f"DOMAIN of non-function {fmt(v)}")


def fcn_apply(f, arg):
    if isinstance(f, tuple):
        if type(arg) is int and 1 <= arg <= len(f):
            return f[arg - 1]
        raise EvalError(f"sequence index {fmt(arg)} out of range for {fmt(f)}")
    if isinstance(f, Fcn):
        try:
            return f.d[arg]
        except KeyError:
            raise EvalError(f"function applied outside its domain: {fmt(f)}[{fmt(arg)}]")
    if isinstance(f, LazyFcn):
        return f.apply(arg)
    raise EvalError(f"applying non-function {fmt(f)}")


def fcn_items(f):
    if isinstance(f, tuple):
        return [(i + 1, x) for i, x in enumerate(f)]
    if isinstance(f, LazyFcn):
        f = f.force()
        return fcn_items(f)
    return list(f.d.items())


class LazyFcn:
    """Recursive function definition f[x \\in S] == e; forced on demand."""

    def __init__(self, domain_thunk, apply_fn):
        self._dom = domain_thunk
        self._apply = apply_fn
        self._memo = {}
        self._forced = None

    def domain(self):
        return self._dom()

    def apply(self, arg):
        m = self._memo
        if arg in m:
            return m[arg]
        v = self._apply(arg)
        m[arg] = v
        return v

    def force(self):
        if self._forced is None:
            dom = to_finite(self._dom())
            self._forced = mk_fcn({k: self.apply(k) for k in sorted_vals(dom)})
        return self._forced


# --------------------------------------------------------------------------
# lazy sets
class LazySet:
    finite = False

    def contains(self, v):
        raise NotImplementedError

    def enumerate(self):
        raise EvalError(f"cannot enumerate {self!r}")


class SetNat(LazySet):
    def contains(self, v):
        return type(v) is int and v >= 0

    def __repr__(self):
        return "Nat"


class SetInt(LazySet):
    def contains(self, v):
        return type(v) is int

    def __repr__(self):
        return "Int"


class SetString(LazySet):
    def contains(self, v):
        return isinstance(v, str)

    def __repr__(self):
        return "STRING"


class SetSeq(LazySet):
    def __init__(self, s):
        self.s = s

    def contains(self, v):
        if not isinstance(v, tuple):
            return False
        return all(set_contains(self.s, x) for x in v)

    def __repr__(self):
        return f"Seq({fmt(self.s)})"


class SetSubset(LazySet):
    finite = True

    def __init__(self, s):
        self.s = s

    def contains(self, v):
        if not isinstance(v, frozenset):
            if isinstance(v, LazySet):
                v = to_finite(v)
            else:
                return False
        return all(set_contains(self.s, x) for x in v)

    def enumerate(self):
        base = sorted_vals(to_finite(self.s))
        for r in range(len(base) + 1):
            for c in itertools.combinations(base, r):
                yield frozenset(c)

    def __repr__(self):
        return f"SUBSET {fmt(self.s)}"


class SetFuncs(LazySet):
    finite = True

    def __init__(self, dom, rng):
        self.dom = dom
        self.rng = rng

    def contains(self, v):
        if not is_fcn_like(v) and not isinstance(v, LazyFcn):
            return False
        if fcn_domain(v) != to_finite(self.dom):
            return False
        return all(set_contains(self.rng, x) for _, x in fcn_items(v))

    def enumerate(self):
        dom = sorted_vals(to_finite(self.dom))
        rng = sorted_vals(to_finite(self.rng))
        for combo in itertools.product(rng, repeat=len(dom)):
            yield mk_fcn(dict(zip(dom, combo)))

    def __repr__(self):
        return f"[{fmt(self.dom)} -> {fmt(self.rng)}]"


class SetPFuncs(LazySet):
    """UNION {[d -> R] : d \\in SUBSET D}: the functions from a subset of D into R (partial functions)."""
    finite = True

    def __init__(self, dom, rng):
        self.dom = dom
        self.rng = rng

    def contains(self, v):
        if isinstance(v, LazyFcn):
            v = v.force()
        if not is_fcn_like(v):
            return False
        return all(set_contains(self.dom, k) and set_contains(self.rng, x) for k, x in fcn_items(v))

    def enumerate(self):
        for d in SetSubset(self.dom).enumerate():
            yield from SetFuncs(d, self.rng).enumerate()

    def __repr__(self):
        return f"UNION {{[d -> {fmt(self.rng)}] : d \\in SUBSET {fmt(self.dom)}}}"


class SetBSeq(LazySet):
    """UNION {[1..k -> S] : k \\in 0..n}: the sequences over S of length at most n."""
    finite = True

    def __init__(self, s, n):
        self.s = s
        self.n = n

    def contains(self, v):
        if isinstance(v, Fcn) and not v.d:
            v = ()
        return isinstance(v, tuple) and len(v) <= self.n and all(set_contains(self.s, x) for x in v)

    def enumerate(self):
        base = sorted_vals(to_finite(self.s))
        for k in range(self.n + 1):
            for c in itertools.product(base, repeat=k):
                yield tuple(c)

    def __repr__(self):
        return f"UNION {{[1..k -> {fmt(self.s)}] : k \\in 0..{self.n}}}"


class SetRecs(LazySet):
    finite = True

    def __init__(self, fields):
        self.fields = fields  # list of (name, set)

    def contains(self, v):
        if not isinstance(v, Fcn):
            return False
        if set(v.d.keys()) != {f for f, _ in self.fields}:
            return False
        return all(set_contains(s, v.d[f]) for f, s in self.fields)

    def enumerate(self):
        names = [f for f, _ in self.fields]
        sets = [sorted_vals(to_finite(s)) for _, s in self.fields]
        for combo in itertools.product(*sets):
            yield Fcn(dict(zip(names, combo)))

    def __repr__(self):
        return "[" + ", ".join(f"{f}: {fmt(s)}" for f, s in self.fields) + "]"


class SetTimes(LazySet):
    finite = True

    def __init__(self, sets):
        self.sets = sets

    def contains(self, v):
        return isinstance(v, tuple) and len(v) == len(self.sets) and \
            all(set_contains(s, x) for s, x in zip(self.sets, v))

    def enumerate(self):
        sets = [sorted_vals(to_finite(s)) for s in self.sets]
        for combo in itertools.product(*sets):
            yield tuple(combo)

    def __repr__(self):
        return " \\X ".join(fmt(s) for s in self.sets)


class SetUnionLazy(LazySet):
    """Union where at least one side is not enumerable (e.g. Nat \\cup {-1})."""

    def __init__(self, a, b):
        self.a = a
        self.b = b
        self.finite = is_enumerable(a) and is_enumerable(b)

    def contains(self, v):
        return set_contains(self.a, v) or set_contains(self.b, v)

    def enumerate(self):
        seen = set()
        for s in (self.a, self.b):
            for x in set_iter(s):
                if x not in seen:
                    seen.add(x)
                    yield x

    def __repr__(self):
        return f"({fmt(self.a)} \\cup {fmt(self.b)})"


def is_set(v):
    return isinstance(v, (frozenset, LazySet))


def is_enumerable(s):
    return isinstance(s, frozenset) or (isinstance(s, LazySet) and s.finite)


def set_contains(s, v):
    if isinstance(s, frozenset):
        try:
            return v in s
        except TypeError:
            return False
    if isinstance(s, LazySet):
        return s.contains(v)
    raise EvalError(f"\\in applied to non-set {fmt(s)}")


def to_finite(s):
    if isinstance(s, frozenset):
        return s
    if isinstance(s, LazySet):
        if not s.finite:
            raise EvalError(f"cannot enumerate infinite set {s!r}")
        return frozenset(s.enumerate())
    raise EvalError(f"expected a set, got {fmt(s)}")


def set_iter(s):
    """Iterate a set in the canonical (TLC-normalised-like) order."""
    if isinstance(s, frozenset):
        return iter(sorted_vals(s))
    if isinstance(s, LazySet):
        if not s.finite:
            raise EvalError(f"cannot enumerate infinite set {s!r}")
        return s.enumerate()
    raise EvalError(f"expected a set, got {fmt(s)}")


# --------------------------------------------------------------------------
# total order used for deterministic enumeration / CHOOSE / printing
def vkey(v):
    t = type(v)
    if t is bool:
        return (0, v)
    if t is int:
        return (1, v)
    if t is str:
        return (2, v)
    if t is ModelValue:
        return (3, v.name)
    if t is tuple:
        return (4, len(v), tuple(vkey(x) for x in v))
    if t is frozenset:
        return (5, len(v), tuple(vkey(x) for x in sorted_vals(v)))
    if t is Fcn:
        items = sorted(((vkey(k), vkey(x)) for k, x in v.d.items()))
        return (6, len(items), tuple(items))
    if isinstance(v, LazySet):
        return vkey(to_finite(v))
    if isinstance(v, LazyFcn):
        return vkey(v.force())
    return (9, repr(v))


_sorted_cache: dict = {}


def sorted_vals(s):
    if isinstance(s, frozenset):
        if len(s) <= 2 and any(type(x) is bool for x in s):
            return tuple(sorted(s))  # never cache: frozenset({0,1}) == frozenset({FALSE,TRUE}) in Python
        r = _sorted_cache.get(s)
        if r is None:
            try:
                # fast path: homogeneous ints / strs
                r = tuple(sorted(s)) if s and all(type(x) is int for x in s) else tuple(sorted(s, key=vkey))
            except TypeError:
                r = tuple(sorted(s, key=vkey))
            if len(_sorted_cache) > 200000:
                _sorted_cache.clear()
            _sorted_cache[s] = r
        return r
    return tuple(sorted(s, key=vkey))


def values_equal(a, b):
    if isinstance(a, LazyFcn):
        a = a.force()
    if isinstance(b, LazyFcn):
        b = b.force()
    if isinstance(a, LazySet) or isinstance(b, LazySet):
        if isinstance(a, LazySet) and isinstance(b, LazySet) and not (a.finite and b.finite):
            return repr(a) == repr(b)
        if not (is_set(a) and is_set(b)):
            return False
        if not is_enumerable(a) or not is_enumerable(b):
            return False
        return to_finite(a) == to_finite(b)
    ta, tb = type(a), type(b)
    if (ta is bool) != (tb is bool):
        return False
    return a == b


# --------------------------------------------------------------------------
# TLC-style printing (format pinned by README.md:270-311 and p-manual p.39)
def fmt(v):
    t = type(v)
    if t is bool:
        return "TRUE" if v else "FALSE"
    if t is int:
        return str(v)
    if t is str:
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if t is ModelValue:
        return v.name
    if t is tuple:
        return "<<" + ", ".join(fmt(x) for x in v) + ">>"
    if t is frozenset:
        return "{" + ", ".join(fmt(x) for x in sorted_vals(v)) + "}"
    if t is Fcn:
        keys = sorted(v.d.keys(), key=vkey)
        if keys and all(isinstance(k, str) for k in keys):
            return "[" + ", ".join(f"{k} |-> {fmt(v.d[k])}" for k in keys) + "]"
        return "(" + " @@ ".join(f"{fmt(k)} :> {fmt(v.d[k])}" for k in keys) + ")"
    if isinstance(v, LazyFcn):
        return fmt(v.force())
    if isinstance(v, LazySet):
        return repr(v)
    return repr(v)


# --------------------------------------------------------------------------
# SYMMETRY support: permutations of model values applied to arbitrary values
def permute_value(v, perm: dict):
    """Apply a permutation of model values (dict ModelValue -> ModelValue) to a value."""
    t = type(v)
    if t is ModelValue:
        return perm.get(v, v)
    if t is tuple:
        return tuple(permute_value(x, perm) for x in v)
    if t is frozenset:
        return frozenset(permute_value(x, perm) for x in v)
    if t is Fcn:
        return mk_fcn({permute_value(k, perm): permute_value(x, perm) for k, x in v.d.items()})
    if isinstance(v, LazyFcn):
        return permute_value(v.force(), perm)
    if isinstance(v, LazySet):
        return permute_value(to_finite(v), perm)
    return v


def permutation_group(perms):
    """Close a set of permutations (dicts) under composition; returns the non-identity elements."""
    dom = sorted({k for p in perms for k in p}, key=lambda m: m.name)
    if not dom:
        return []

    def norm(p):
        return tuple(p.get(k, k) for k in dom)

    ident = tuple(dom)
    gens = {norm(p) for p in perms if norm(p) != ident}
    group = {ident}
    frontier = [ident]
    while frontier:
        nxt = []
        for g in frontier:
            gd = dict(zip(dom, g))
            for h in gens:
                hd = dict(zip(dom, h))
                comp = tuple(hd[gd[k]] for k in dom)     # h o g
                if comp not in group:
                    group.add(comp)
                    nxt.append(comp)
        frontier = nxt
        if len(group) > 50000:
            raise EvalError("symmetry group too large")
    return [dict(zip(dom, g)) for g in sorted(group, key=lambda g: tuple(x.name for x in g)) if g != ident]
