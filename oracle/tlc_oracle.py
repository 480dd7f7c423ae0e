"""ORACLE (test infrastructure, not product): CPU restatement of TLC's explicit-state
breadth-first model-checking loop, the hot path named by BASELINE.json.

The algorithm itself lives in tlaplus/tlaplus (tla2tools.jar: tlc2.TLC /
tlc2.tool.ModelChecker + Worker), which is NOT under /root/reference and whose
version the reference does not pin (Makefile:7 just runs `tlc *tla`).  This file
restates the published algorithm (p-manual.pdf §4 "found all
reachable states using a breadth-first search"):

    enumerate Init; for each state in FIFO order: enumerate Next successors in
    syntactic/ascending order; generated++ per successor; fingerprint-set insert;
    invariants checked on first sight; deadlock = no successor; stop at first error.

Pinned against the only known-answer transcript the reference holds for this path,
README.md:267-321 (9097 generated / 6164 distinct / 999 on queue / depth 7 and the
6-state trace) -- see tests/test_oracle_golden.py.  For every other BASELINE config
the reference holds no counts ("parity unpinned" at the TLC boundary);
there the oracle is the reference point for the CUDA engine.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
It evaluates the AST directly with Python sets/tuples (tla_rust_b200.front.eval is
the shared *front end*: parser + value semantics); it shares no code with the
bytecode compiler or the CUDA engine.
"""
from __future__ import annotations

import time
from collections import deque

from tla_rust_b200.front.eval import AssertFailure, Fr
from tla_rust_b200.front.values import EvalError
from tla_rust_b200.front.report import (CheckResult, OK, INVARIANT, ASSERT, DEADLOCK, EVAL_ERROR, PROPERTY)


class Oracle:
    def __init__(self, model):
        self.m = model
        self.ev = model.ev
        self.vars = model.vars
        self.group = model.symmetry_group()

    def key(self, st):
        return tuple(st[v] for v in self.vars)

    def canon(self, st):
        """Representative of st's orbit under the SYMMETRY group: the permuted state that is least in the
        canonical value order (any fixed choice yields the same orbit count)."""
        if not self.group:
            return st
        from tla_rust_b200.front.values import permute_value, vkey
        best, bk = st, tuple(vkey(st[v]) for v in self.vars)
        for p in self.group:
            c = {v: permute_value(st[v], p) for v in self.vars}
            ck = tuple(vkey(c[v]) for v in self.vars)
            if ck < bk:
                best, bk = c, ck
        return best

    def successors(self, st):
        """Yield (next-state dict, action label) in TLC's enumeration order."""
        m = self.m
        nvars = len(self.vars)
        for asg, act in self.ev.solve(m.next_node, {}, m.next_ctx, st, {}, "next"):
            if len(asg) != nvars:
                missing = [v for v in self.vars if v not in asg]
                raise EvalError(f"action does not assign {missing}")
            yield asg, act

    def check_invariants(self, st):
        for nm, node, ctx in self.m.invariants:
            if self.ev.eval(node, {}, Fr(ctx, st, None)) is not True:
                return nm
        return None

    def check_refinements(self, s, t):
        """[Next2]_v2 on the transition s -> t for every refinement PROPERTY (safety part)."""
        from tla_rust_b200.front.values import values_equal
        for nm, _, nxt, sub, ctx in self.m.refinements:
            if nxt is None:
                continue
            fr = Fr(ctx, s, t)
            if self.ev.eval(nxt, {}, fr) is True:
                continue
            if values_equal(self.ev.eval(sub, {}, fr), self.ev.eval(sub, {}, Fr(ctx, t, None))):
                continue
            return nm
        return None

    def in_model(self, st):
        for nm, node, ctx in self.m.constraints:
            if self.ev.eval(node, {}, Fr(ctx, st, None)) is not True:
                return False
        return True

    def in_actions(self, s, t):
        for nm, node, ctx in self.m.action_constraints:
            if self.ev.eval(node, {}, Fr(ctx, s, t)) is not True:
                return False
        return True

    def run(self, max_states=None, progress=None) -> CheckResult:
        t0 = time.time()
        res = CheckResult()
        seen = {}          # key -> index
        parent = []        # index -> (parent index or -1, act)
        states = []        # index -> state dict
        level = []
        queue = deque()
        maxlevel = 0

        def trace_to(idx, extra=None):
            chain = []
            while idx >= 0:
                p, act = parent[idx]
                chain.append((states[idx], act))
                idx = p
            chain.reverse()
            if extra is not None:
                chain.append(extra)
            return chain

        def finish(verdict):
            res.verdict = verdict
            res.distinct = len(states)
            res.queue = len(queue)
            res.depth = maxlevel
            res.seconds = time.time() - t0
            return res

        try:
            inits = self.m.initial_states()
        except AssertFailure as af:
            res.error_text = af.msg
            return finish(ASSERT)
        for st in inits:
            res.generated += 1
            st = self.canon(st)
            k = self.key(st)
            if k in seen:
                continue
            bad = self.check_invariants(st)
            idx = len(states)
            seen[k] = idx
            states.append(st)
            parent.append((-1, None))
            level.append(1)
            maxlevel = 1
            pbad = self.m.check_refinement_init(st)
            if pbad is not None:
                res.invariant = pbad
                res.error_text = f"Property {pbad} is violated by the initial state"
                res.trace = trace_to(idx)
                return finish(PROPERTY)
            if bad is not None:
                res.invariant = bad
                res.trace = trace_to(idx)
                return finish(INVARIANT)
            if self.in_model(st):
                queue.append(idx)
        res.init_states = len(states)
        deadlock = self.m.check_deadlock
        while queue:
            cur = queue.popleft()
            st = states[cur]
            lv = level[cur]
            nsucc = 0
            try:
                for t, act in self.successors(st):
                    nsucc += 1
                    res.generated += 1
                    if self.m.refinements:
                        pbad = self.check_refinements(st, t)
                        if pbad is not None:
                            res.invariant = pbad
                            res.trace = trace_to(cur, (t, act))
                            return finish(PROPERTY)
                    in_model = self.in_model(t) and self.in_actions(st, t)
                    t = self.canon(t)
                    k = self.key(t)
                    is_seen = False
                    idx = -1
                    if in_model:
                        idx = seen.get(k, -1)
                        is_seen = idx >= 0
                        if not is_seen:
                            idx = len(states)
                            seen[k] = idx
                            states.append(t)
                            parent.append((cur, act))
                            level.append(lv + 1)
                            if lv + 1 > maxlevel:
                                maxlevel = lv + 1
                            queue.append(idx)
                    if not is_seen:
                        bad = self.check_invariants(t)
                        if bad is not None:
                            res.invariant = bad
                            res.trace = trace_to(idx) if idx >= 0 else trace_to(cur, (t, act))
                            return finish(INVARIANT)
            except AssertFailure as af:
                res.error_text = af.msg
                res.trace = trace_to(cur)
                if af.node is not None:
                    res.extra["assert_loc"] = af.node.loc()
                if lv + 1 > maxlevel:
                    maxlevel = lv + 1
                return finish(ASSERT)
            except EvalError as ex:
                res.error_text = str(ex)
                res.trace = trace_to(cur)
                return finish(EVAL_ERROR)
            if nsucc == 0 and deadlock:
                res.trace = trace_to(cur)
                return finish(DEADLOCK)
            if max_states is not None and len(states) >= max_states:
                res.extra["truncated"] = True
                return finish(OK)
            if progress and len(states) % progress == 0:
                print(f"Progress({lv}): {res.generated} generated, {len(states)} distinct, {len(queue)} on queue",
                      flush=True)
        self.states = states
        self.level = level
        return finish(OK)
