"""Finite types for the fixed-width state-vector encoding.

Every TLA+ value that reaches the engine has a static type T with
  size(T)   number of 32-bit words of its *frame representation* (unpacked, one
            word per scalar -- what the bytecode VM computes on), and
  bits(T)   its *packed* width inside a stored state vector (what lives in HBM).

Representations (frame words):
  TInt(lo,hi)   1 word, the integer itself (packed biased by lo; lo=None: raw 32 bit)
  TBool         1 word 0/1
  TAtom(atoms)  1 word, global atom id (strings and model values share one id space)
  TRec / TUnion [tag?] + one slot per field name (union of all alternatives' fields)
  TTuple        concatenation
  TFun(keys,T)  |keys| consecutive elements, keys in canonical order
  TSet(E)       bitset over the enumeration of E: ceil(card(E)/32) words
  TSeq(E,cap)   length word + cap elements (unused slots zero => canonical)
  TPFun(keys,T) partial function over a constant key list: per key a presence word + element (absent => zero)
  TSparse(K,V,cap)  bounded sparse container: length word + cap entries (K [, V]) kept sorted by the signed
                word-wise order of the key words, no duplicate keys, unused entries zero => canonical.
                V = None: a set of K whose universe is too large for a bitset; otherwise a function with a
                dynamic domain (raft's message bag, raft.tla:35).
Enumerable types (everything except unbounded ints) have card(T), an ordinal
bijection ord/unord (Horner / mixed radix), used for set universes and function keys.
"""
from __future__ import annotations

import itertools

from ..front.values import (ModelValue, Fcn, LazySet, SetNat, SetInt, SetString, SetSeq, SetSubset, SetFuncs,
                            SetRecs, SetTimes, SetUnionLazy, SetPFuncs, SetBSeq, mk_fcn, sorted_vals, fmt, vkey,
                            LazyFcn, fcn_items, is_fcn_like)

MAX_SET_BITS = 8192
ATOM_ALT = "\x00atom"     # pseudo-field of the record alternative that holds a bare atom (atom | record unions)
SPARSE_CAP = 8          # default capacity of sparse containers (overridden per variable by a Cardinality bound)
INT32_MIN = -(1 << 31)
INT32_MAX = (1 << 31) - 1


class TypeErr(Exception):
    pass


class Atoms:
    """Global atom table: strings and model values -> small ids (id 0 is reserved)."""

    def __init__(self):
        self.ids = {}
        self.vals = [None]

    def id(self, v):
        k = ("s", v) if isinstance(v, str) else ("m", v.name)
        i = self.ids.get(k)
        if i is None:
            i = len(self.vals)
            self.ids[k] = i
            self.vals.append(v)
        return i

    def val(self, i):
        return self.vals[i]

    def width(self):
        return max(1, (len(self.vals) - 1).bit_length()) if len(self.vals) > 1 else 1


def is_atom(v):
    return isinstance(v, (str, ModelValue))


class T:
    size = 1
    scalar = False

    def card(self):
        raise TypeErr(f"type {self} is not enumerable")

    def __eq__(self, o):
        return type(self) is type(o) and self.key() == o.key()

    def __hash__(self):
        return hash((type(self).__name__, self.key()))

    def __repr__(self):
        return f"{type(self).__name__}{self.key()}"


class TInt(T):
    scalar = True

    def __init__(self, lo=None, hi=None):
        self.lo, self.hi = lo, hi

    def key(self):
        return (self.lo, self.hi)

    def card(self):
        if self.lo is None:
            raise TypeErr("unbounded integer type is not enumerable (bound it with a TypeOK)")
        return self.hi - self.lo + 1


class TBool(T):
    scalar = True

    def key(self):
        return ()

    def card(self):
        return 2


class TAtom(T):
    scalar = True

    def __init__(self, atoms):
        self.atoms = tuple(atoms)  # python values, canonical (vkey) order

    def key(self):
        return tuple(vkey(a) for a in self.atoms)

    def card(self):
        return len(self.atoms)


class TRec(T):
    """Record or tagged union of records.

    alts      sorted list of field-name tuples (one per alternative)
    alt_types per alternative: dict field -> T (used for cardinality / ordinals)
    fields    dict field -> T: one *slot* per distinct field name, typed with the join over
              the alternatives that have it (frame representation and packed layout)."""

    def __init__(self, alts, fields=None):
        per = {}
        for a in alts:
            if isinstance(a, dict):
                names = tuple(sorted(a))
                types = dict(a)
            else:
                names = tuple(sorted(a))
                types = {f: fields[f] for f in names}
            if names in per:
                per[names] = {f: join(per[names][f], types[f]) for f in names}
            else:
                per[names] = types
        self.alts = sorted(per)
        self.alt_types = [per[a] for a in self.alts]
        slot = {}
        for a, ts in zip(self.alts, self.alt_types):
            for f in a:
                slot[f] = join(slot.get(f), ts[f])
        self.fields = slot
        self.fnames = sorted(slot)
        self.tagged = len(self.alts) > 1
        off = 1 if self.tagged else 0
        self.off = {}
        for f in self.fnames:
            self.off[f] = off
            off += self.fields[f].size
        self.size = off

    def key(self):
        return tuple((a, tuple((f, ts[f]) for f in a)) for a, ts in zip(self.alts, self.alt_types))

    def alt_dicts(self):
        return [dict(ts) for ts in self.alt_types]

    def alt_card(self, i):
        c = 1
        for f in self.alts[i]:
            c *= self.alt_types[i][f].card()
        return c

    def card(self):
        return sum(self.alt_card(i) for i in range(len(self.alts)))

    def alt_index(self, names):
        t = tuple(sorted(names))
        try:
            return self.alts.index(t)
        except ValueError:
            return -1

    def alt_base(self, i):
        return sum(self.alt_card(j) for j in range(i))


class TTuple(T):
    def __init__(self, elems):
        self.elems = list(elems)
        self.size = sum(e.size for e in self.elems)
        self.offs = list(itertools.accumulate([0] + [e.size for e in self.elems]))[:-1]

    def key(self):
        return tuple(self.elems)

    def card(self):
        c = 1
        for e in self.elems:
            c *= e.card()
        return c


class TFun(T):
    def __init__(self, keys, elem, keyt=None):
        self.keys = tuple(keys)        # python values in canonical order
        self.elem = elem
        self.size = len(self.keys) * elem.size
        self.kindex = {k: i for i, k in enumerate(self.keys)}
        self.keyt = keyt

    def key(self):
        return (tuple(vkey(k) for k in self.keys), self.elem)

    def card(self):
        return self.elem.card() ** len(self.keys)


def has_dynamic(t) -> bool:
    """True if t contains a sequence / partial function / sparse container (no ordinal encoding)."""
    if isinstance(t, (TSeq, TPFun, TSparse)):
        return True
    if isinstance(t, TRec):
        return any(has_dynamic(x) for x in t.fields.values())
    if isinstance(t, TTuple):
        return any(has_dynamic(x) for x in t.elems)
    if isinstance(t, TFun):
        return has_dynamic(t.elem)
    return False


class TSet(T):
    def __init__(self, elem):
        self.elem = elem
        if has_dynamic(elem):
            raise TypeErr(f"no bitset universe for element type {elem}")
        n = elem.card()
        if n > MAX_SET_BITS:
            raise TypeErr(f"set universe of {n} elements exceeds the {MAX_SET_BITS}-bit limit ({elem})")
        self.nbits = n
        self.size = max(1, (n + 31) // 32)

    def key(self):
        return (self.elem,)

    def card(self):
        if self.nbits > 30:
            raise TypeErr("set of sets too large to enumerate")
        return 1 << self.nbits


class TSeq(T):
    def __init__(self, elem, cap):
        self.elem = elem
        self.cap = cap
        self.size = 1 + cap * elem.size

    def key(self):
        return (self.elem, self.cap)

    def card(self):
        c = self.elem.card()
        return sum(c ** k for k in range(self.cap + 1))


class TPFun(T):
    def __init__(self, keys, elem):
        self.keys = tuple(keys)
        self.elem = elem
        self.stride = 1 + elem.size
        self.size = len(self.keys) * self.stride
        self.kindex = {k: i for i, k in enumerate(self.keys)}

    def key(self):
        return (tuple(vkey(k) for k in self.keys), self.elem)

    def card(self):
        return (self.elem.card() + 1) ** len(self.keys)


class TSparse(T):
    def __init__(self, kt, vt, cap):
        self.kt, self.vt, self.cap = kt, vt, cap
        self.keyw = kt.size
        self.stride = kt.size + (vt.size if vt is not None else 0)
        if self.stride > 127:
            raise TypeErr(f"sparse container entry of {self.stride} words exceeds the 127-word limit")
        self.size = 1 + cap * self.stride

    def key(self):
        return (self.kt, self.vt, self.cap)

    def card(self):
        raise TypeErr("sparse containers are not enumerable")


class TBottom(T):
    """Type of an empty set's elements / unknown; joins with anything."""
    size = 0

    def key(self):
        return ()

    def card(self):
        return 0


# ---------------------------------------------------------------------------
def join(a: T, b: T) -> T:
    if a is None:
        return b
    if b is None:
        return a
    if isinstance(a, TBottom):
        return b
    if isinstance(b, TBottom):
        return a
    if a == b:
        return a
    ta, tb = type(a), type(b)
    if ta is TInt and tb is TInt:
        if a.lo is None or b.lo is None:
            return TInt()
        return TInt(min(a.lo, b.lo), max(a.hi, b.hi))
    if ta is TAtom and tb is TAtom:
        return TAtom(sorted(set(a.atoms) | set(b.atoms), key=vkey))
    if ta is TBool and tb is TBool:
        return a
    if ta is TSet and tb is TSet:
        return TSet(join(a.elem, b.elem))
    if ta is TRec and tb is TRec:
        return TRec(a.alt_dicts() + b.alt_dicts())
    # a slot that holds either a record or a bare atom (InternalMemory.tla:12: buf \in [Proc -> MReq \cup Val \cup
    # {NoVal}]): the atom becomes one more alternative of the tagged union
    if ta is TRec and tb is TAtom:
        return TRec(a.alt_dicts() + [{ATOM_ALT: b}])
    if ta is TAtom and tb is TRec:
        return join(b, a)
    if ta is TTuple and tb is TTuple and len(a.elems) == len(b.elems):
        return TTuple([join(x, y) for x, y in zip(a.elems, b.elems)])
    if ta is TFun and tb is TFun and a.keys == b.keys:
        return TFun(a.keys, join(a.elem, b.elem))
    if ta is TSeq and tb is TSeq:
        return TSeq(join(a.elem, b.elem), max(a.cap, b.cap))
    if ta is TTuple and tb is TTuple:
        e = TBottom()
        for x in a.elems + b.elems:
            e = join(e, x)
        return TSeq(e, max(len(a.elems), len(b.elems)))
    if ta is TPFun and tb is TPFun and a.keys == b.keys:
        return TPFun(a.keys, join(a.elem, b.elem))
    if ta is TSparse and tb is TSparse and (a.vt is None) == (b.vt is None):
        return TSparse(join(a.kt, b.kt), None if a.vt is None else join(a.vt, b.vt), max(a.cap, b.cap))
    if ta is TSet and tb is TSparse and b.vt is None:
        return TSparse(join(a.elem, b.kt), None, b.cap)
    if ta is TSparse and tb is TSet and a.vt is None:
        return TSparse(join(a.kt, b.elem), None, a.cap)
    if ta is TSeq and tb is TTuple:
        return join(b, a)
    if ta is TTuple and tb is TSeq:
        e = b.elem
        for x in a.elems:
            e = join(e, x)
        return TSeq(e, max(b.cap, len(a.elems)))
    raise TypeErr(f"cannot unify types {a} and {b}")


def type_of_value(v, seq_cap=None) -> T:
    """Most specific type of a concrete value (ints become singleton ranges)."""
    if isinstance(v, LazyFcn):
        v = v.force()
    if type(v) is bool:
        return TBool()
    if type(v) is int:
        return TInt(v, v)
    if is_atom(v):
        return TAtom([v])
    if isinstance(v, tuple):
        return TTuple([type_of_value(x, seq_cap) for x in v])
    if isinstance(v, frozenset):
        e = TBottom()
        for x in v:
            e = join(e, type_of_value(x, seq_cap))
        return TSet(e) if not isinstance(e, TBottom) else TSet(TBottom())
    if isinstance(v, Fcn):
        keys = sorted(v.d.keys(), key=vkey)
        if all(isinstance(k, str) for k in keys):
            return TRec([keys], {k: type_of_value(v.d[k], seq_cap) for k in keys})
        e = TBottom()
        for k in keys:
            e = join(e, type_of_value(v.d[k], seq_cap))
        return TFun(keys, e)
    if isinstance(v, LazySet):
        return TSet(type_of_set(v, seq_cap))
    raise TypeErr(f"cannot type value {fmt(v)}")


def type_of_set(s, seq_cap=None) -> T:
    """Element type of a set value (lazy sets keep their structure: [S -> T], SUBSET S, ...)."""
    if isinstance(s, frozenset):
        e = TBottom()
        for x in s:
            e = join(e, type_of_value(x, seq_cap))
        return e
    if isinstance(s, SetFuncs):
        dom = sorted_vals(frozenset(s.dom.enumerate()) if isinstance(s.dom, LazySet) else s.dom)
        rng = type_of_set(s.rng, seq_cap)
        n = len(dom)
        if n > 0 and all(type(k) is int for k in dom) and dom[0] == 1 and dom[-1] == n:
            return TTuple([rng] * n)
        return TFun(dom, rng)
    if isinstance(s, SetSubset):
        et = type_of_set(s.s, seq_cap)
        try:
            return TSet(et)
        except TypeErr:
            return TSparse(et, None, SPARSE_CAP)     # universe too large (or not enumerable) for a bitset
    if isinstance(s, SetPFuncs):
        return TPFun(sorted_vals(to_finite_keys(s.dom)), type_of_set(s.rng, seq_cap))
    if isinstance(s, SetBSeq):
        return TSeq(type_of_set(s.s, seq_cap), s.n)
    if isinstance(s, SetRecs):
        names = [f for f, _ in s.fields]
        return TRec([names], {f: type_of_set(x, seq_cap) for f, x in s.fields})
    if isinstance(s, SetTimes):
        return TTuple([type_of_set(x, seq_cap) for x in s.sets])
    if isinstance(s, SetUnionLazy):
        return join(type_of_set(s.a, seq_cap), type_of_set(s.b, seq_cap))
    if isinstance(s, SetSeq):
        if seq_cap is None:
            raise TypeErr("Seq(S) needs a sequence capacity (option seq_cap)")
        return TSeq(type_of_set(s.s, seq_cap), seq_cap)
    if isinstance(s, (SetNat, SetInt)):
        return TInt()
    if isinstance(s, SetString):
        raise TypeErr("STRING is not a finite type")
    raise TypeErr(f"cannot derive a type from set {s!r}")


def to_finite_keys(d):
    return frozenset(d.enumerate()) if isinstance(d, LazySet) else d


def subset_type(et: T) -> T:
    """Type of a subset of a set with element type et: a bitset when the universe is small, else sparse."""
    try:
        return TSet(et)
    except TypeErr:
        return TSparse(et, None, SPARSE_CAP)


def _zero_safe(slots):
    """Re-bias packed slots so that the frame value 0 is representable (slots that can be inactive)."""
    out = []
    for off, w, b in slots:
        if w > 0 and b > 0:
            hi = b + (1 << w) - 1
            out.append((off, max(1, hi.bit_length()), 0))
        else:
            out.append((off, w, b))
    return out


def widen_init(t: T, all_atoms) -> T:
    """Widen a type inferred from initial values only: ints -> int32, atoms -> every atom."""
    if isinstance(t, TInt):
        return TInt()
    if isinstance(t, TAtom):
        return TAtom(all_atoms)
    if isinstance(t, TBool):
        return t
    if isinstance(t, TTuple):
        return TTuple([widen_init(e, all_atoms) for e in t.elems])
    if isinstance(t, TFun):
        return TFun(t.keys, widen_init(t.elem, all_atoms))
    if isinstance(t, TRec):
        return TRec([{f: widen_init(x, all_atoms) for f, x in d.items()} for d in t.alt_dicts()])
    if isinstance(t, TSet):
        if isinstance(t.elem, TBottom):
            raise TypeErr("cannot infer the element type of an initially-empty set; provide a TypeOK")
        e = t.elem
        if isinstance(e, TAtom):
            return TSet(TAtom(all_atoms))
        return t
    if isinstance(t, TSeq):
        return TSeq(widen_init(t.elem, all_atoms), t.cap)
    return t


# ---------------------------------------------------------------------------
# python-side encoding (init states, constants, trace decoding)
class Codec:
    def __init__(self, atoms: Atoms):
        self.atoms = atoms

    def ord_of(self, t: T, v) -> int:
        """Ordinal of value v in the enumeration of t, or -1 if v is outside t."""
        if isinstance(t, TInt):
            if type(v) is not int or t.lo is None or not (t.lo <= v <= t.hi):
                return -1
            return v - t.lo
        if isinstance(t, TBool):
            return int(v) if type(v) is bool else -1
        if isinstance(t, TAtom):
            try:
                return t.atoms.index(v) if is_atom(v) else -1
            except ValueError:
                return -1
        if isinstance(t, TRec):
            if is_atom(v) and ATOM_ALT in t.fields:
                v = Fcn({ATOM_ALT: v})
            if not isinstance(v, Fcn):
                return -1
            ai = t.alt_index(v.d.keys())
            if ai < 0:
                return -1
            acc = 0
            for f in t.alts[ai]:
                ft = t.alt_types[ai][f]
                o = self.ord_of(ft, v.d[f])
                if o < 0:
                    return -1
                acc = acc * ft.card() + o
            return t.alt_base(ai) + acc
        if isinstance(t, TTuple):
            if not isinstance(v, tuple) or len(v) != len(t.elems):
                return -1
            acc = 0
            for e, x in zip(t.elems, v):
                o = self.ord_of(e, x)
                if o < 0:
                    return -1
                acc = acc * e.card() + o
            return acc
        if isinstance(t, TSet):
            if isinstance(v, LazySet):
                v = frozenset(v.enumerate())
            if not isinstance(v, frozenset):
                return -1
            m = 0
            for x in v:
                o = self.ord_of(t.elem, x)
                if o < 0:
                    return -1
                m |= 1 << o
            return m
        if isinstance(t, TFun):
            if isinstance(v, tuple):
                d = {i + 1: x for i, x in enumerate(v)}
            elif isinstance(v, Fcn):
                d = v.d
            else:
                return -1
            if set(d.keys()) != set(t.keys):
                return -1
            acc = 0
            for k in t.keys:
                o = self.ord_of(t.elem, d[k])
                if o < 0:
                    return -1
                acc = acc * t.elem.card() + o
            return acc
        raise TypeErr(f"ord_of: unsupported type {t}")

    def enum(self, t: T):
        """All values of t in ordinal order."""
        if isinstance(t, TInt):
            return list(range(t.lo, t.hi + 1))
        if isinstance(t, TBool):
            return [False, True]
        if isinstance(t, TAtom):
            return list(t.atoms)
        if isinstance(t, TRec):
            out = []
            for alt, ts in zip(t.alts, t.alt_types):
                for combo in itertools.product(*[self.enum(ts[f]) for f in alt]):
                    out.append(combo[0] if alt == (ATOM_ALT,) else Fcn(dict(zip(alt, combo))))
            return out
        if isinstance(t, TTuple):
            return [tuple(c) for c in itertools.product(*[self.enum(e) for e in t.elems])]
        if isinstance(t, TSet):
            base = self.enum(t.elem)
            return [frozenset(base[i] for i in range(len(base)) if (m >> i) & 1) for m in range(1 << len(base))]
        if isinstance(t, TFun):
            ev = self.enum(t.elem)
            return [mk_fcn(dict(zip(t.keys, c))) for c in itertools.product(ev, repeat=len(t.keys))]
        raise TypeErr(f"enum: unsupported type {t}")

    def rep(self, t: T, v) -> list:
        """Frame representation (list of int32 words) of value v at type t."""
        if isinstance(v, LazyFcn):
            v = v.force()
        if isinstance(t, TInt):
            if type(v) is not int:
                raise TypeErr(f"expected integer, got {fmt(v)}")
            if t.lo is not None and not (t.lo <= v <= t.hi):
                raise TypeErr(f"integer {v} outside {t.lo}..{t.hi}")
            if not (INT32_MIN <= v <= INT32_MAX):
                raise TypeErr(f"integer {v} outside int32")
            return [v]
        if isinstance(t, TBool):
            if type(v) is not bool:
                raise TypeErr(f"expected BOOLEAN, got {fmt(v)}")
            return [int(v)]
        if isinstance(t, TAtom):
            if not is_atom(v):
                raise TypeErr(f"expected string/model value, got {fmt(v)}")
            return [self.atoms.id(v)]
        if isinstance(t, TRec):
            if is_atom(v) and ATOM_ALT in t.fields:
                v = Fcn({ATOM_ALT: v})
            if not isinstance(v, Fcn):
                raise TypeErr(f"expected record, got {fmt(v)}")
            ai = t.alt_index(v.d.keys())
            if ai < 0:
                raise TypeErr(f"record {fmt(v)} does not fit type {t}")
            out = [0] * t.size
            if t.tagged:
                out[0] = ai
            for f in t.alts[ai]:
                r = self.rep(t.fields[f], v.d[f])
                out[t.off[f]:t.off[f] + len(r)] = r
            return out
        if isinstance(t, TTuple):
            if not isinstance(v, tuple) or len(v) != len(t.elems):
                raise TypeErr(f"expected {len(t.elems)}-tuple, got {fmt(v)}")
            out = []
            for e, x in zip(t.elems, v):
                out += self.rep(e, x)
            return out
        if isinstance(t, TFun):
            if isinstance(v, tuple):
                d = {i + 1: x for i, x in enumerate(v)}
            elif isinstance(v, Fcn):
                d = v.d
            else:
                raise TypeErr(f"expected function, got {fmt(v)}")
            if set(d.keys()) != set(t.keys):
                raise TypeErr(f"function domain {fmt(frozenset(d.keys()))} does not match type")
            out = []
            for k in t.keys:
                out += self.rep(t.elem, d[k])
            return out
        if isinstance(t, TSet):
            if isinstance(v, LazySet):
                v = frozenset(v.enumerate())
            if not isinstance(v, frozenset):
                raise TypeErr(f"expected set, got {fmt(v)}")
            words = [0] * t.size
            for x in v:
                o = self.ord_of(t.elem, x)
                if o < 0:
                    raise TypeErr(f"set element {fmt(x)} outside universe {t.elem}")
                words[o >> 5] |= 1 << (o & 31)
            return [w - (1 << 32) if w >= (1 << 31) else w for w in words]
        if isinstance(t, TSeq):
            if not isinstance(v, tuple):
                raise TypeErr(f"expected sequence, got {fmt(v)}")
            if len(v) > t.cap:
                raise TypeErr(f"sequence longer than capacity {t.cap}")
            out = [len(v)]
            for x in v:
                out += self.rep(t.elem, x)
            out += [0] * (t.size - len(out))
            return out
        if isinstance(t, TPFun):
            if not is_fcn_like(v):
                raise TypeErr(f"expected function, got {fmt(v)}")
            d = dict(fcn_items(v))
            if not set(d) <= set(t.keys):
                raise TypeErr(f"function domain {fmt(frozenset(d))} outside the key set of its type")
            out = []
            for k in t.keys:
                if k in d:
                    out += [1] + self.rep(t.elem, d[k])
                else:
                    out += [0] * t.stride
            return out
        if isinstance(t, TSparse):
            if t.vt is None:
                if isinstance(v, LazySet):
                    v = frozenset(v.enumerate())
                if not isinstance(v, frozenset):
                    raise TypeErr(f"expected set, got {fmt(v)}")
                ents = [self.rep(t.kt, x) for x in v]
            else:
                if not is_fcn_like(v):
                    raise TypeErr(f"expected function, got {fmt(v)}")
                ents = [self.rep(t.kt, k) + self.rep(t.vt, x) for k, x in fcn_items(v)]
            if len(ents) > t.cap:
                raise TypeErr(f"{len(ents)} entries exceed the sparse capacity {t.cap}")
            ents.sort()
            out = [len(ents)]
            for e in ents:
                out += e
            out += [0] * (t.size - len(out))
            return out
        raise TypeErr(f"rep: unsupported type {t}")

    def unrep(self, t: T, w, i=0):
        """Inverse of rep: decode words[i:i+t.size] -> python value."""
        if isinstance(t, TInt):
            return int(w[i])
        if isinstance(t, TBool):
            return bool(w[i])
        if isinstance(t, TAtom):
            return self.atoms.val(int(w[i]))
        if isinstance(t, TRec):
            ai = int(w[i]) if t.tagged else 0
            if t.alts[ai] == (ATOM_ALT,):
                return self.unrep(t.fields[ATOM_ALT], w, i + t.off[ATOM_ALT])
            return Fcn({f: self.unrep(t.fields[f], w, i + t.off[f]) for f in t.alts[ai]})
        if isinstance(t, TTuple):
            return tuple(self.unrep(e, w, i + o) for e, o in zip(t.elems, t.offs))
        if isinstance(t, TFun):
            return mk_fcn({k: self.unrep(t.elem, w, i + j * t.elem.size) for j, k in enumerate(t.keys)})
        if isinstance(t, TSet):
            base = self.enum(t.elem)
            out = []
            for b in range(t.nbits):
                if (int(w[i + (b >> 5)]) >> (b & 31)) & 1:
                    out.append(base[b])
            return frozenset(out)
        if isinstance(t, TSeq):
            n = int(w[i])
            return tuple(self.unrep(t.elem, w, i + 1 + j * t.elem.size) for j in range(n))
        if isinstance(t, TPFun):
            d = {}
            for j, k in enumerate(t.keys):
                if int(w[i + j * t.stride]):
                    d[k] = self.unrep(t.elem, w, i + j * t.stride + 1)
            return mk_fcn(d)
        if isinstance(t, TSparse):
            n = int(w[i])
            if t.vt is None:
                return frozenset(self.unrep(t.kt, w, i + 1 + j * t.stride) for j in range(n))
            return mk_fcn({self.unrep(t.kt, w, i + 1 + j * t.stride): self.unrep(t.vt, w, i + 1 + j * t.stride + t.keyw)
                           for j in range(n)})
        raise TypeErr(f"unrep: unsupported type {t}")

    # -- packed layout -------------------------------------------------------
    def layout(self, t: T, off=0):
        """List of (frame_off, width_bits, bias) scalar slots, in frame order."""
        if isinstance(t, TInt):
            if t.lo is None:
                return [(off, 32, 0)]
            return [(off, max(1, (t.hi - t.lo).bit_length()), t.lo)]
        if isinstance(t, TBool):
            return [(off, 1, 0)]
        if isinstance(t, TAtom):
            return [(off, -1, 0)]   # width resolved once the atom table is final
        if isinstance(t, TRec):
            out = []
            if t.tagged:
                out.append((off, max(1, (len(t.alts) - 1).bit_length()), 0))
            for f in t.fnames:
                sub = self.layout(t.fields[f], off + t.off[f])
                out += _zero_safe(sub) if t.tagged else sub
            return out
        if isinstance(t, TTuple):
            out = []
            for e, o in zip(t.elems, t.offs):
                out += self.layout(e, off + o)
            return out
        if isinstance(t, TFun):
            out = []
            for j in range(len(t.keys)):
                out += self.layout(t.elem, off + j * t.elem.size)
            return out
        if isinstance(t, TSet):
            out = []
            left = t.nbits
            for j in range(t.size):
                out.append((off + j, min(32, left) if left > 0 else 1, 0))
                left -= 32
            return out
        if isinstance(t, TSeq):
            out = [(off, max(1, t.cap.bit_length()), 0)]
            for j in range(t.cap):
                out += _zero_safe(self.layout(t.elem, off + 1 + j * t.elem.size))
            return out
        if isinstance(t, TPFun):
            out = []
            for j in range(len(t.keys)):
                out.append((off + j * t.stride, 1, 0))
                out += _zero_safe(self.layout(t.elem, off + j * t.stride + 1))
            return out
        if isinstance(t, TSparse):
            out = [(off, max(1, t.cap.bit_length()), 0)]
            for j in range(t.cap):
                o = off + 1 + j * t.stride
                out += _zero_safe(self.layout(t.kt, o))
                if t.vt is not None:
                    out += _zero_safe(self.layout(t.vt, o + t.keyw))
            return out
        raise TypeErr(f"layout: unsupported type {t}")
