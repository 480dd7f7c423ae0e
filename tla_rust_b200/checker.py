"""Host driver pieces: Model -> bytecode, state encode/decode, engine result -> TLC report.

This is the `tlc` half of the reference's CLI contract (Makefile:6-7): for each FILE.tla
load FILE.cfg, check, print in TLC's format (README.md:267-321), non-zero exit on error.
"""
from __future__ import annotations

import numpy as np

from .front.spec import Model
from .front.report import CheckResult, OK, INVARIANT, ASSERT, DEADLOCK, EVAL_ERROR, PROPERTY
from .compile.lower import Lowering, CompiledModel, CompileError, CompileBudget
from .compile.bytecode import TRAP_NAMES


def compile_model(model: Model, init_states=None, seq_cap=None, type_hint=None, subroutines=False) -> CompiledModel:
    if init_states is None:
        init_states = model.initial_states()
    lw = Lowering(model, seq_cap=seq_cap, type_hint=type_hint)
    if subroutines:
        lw.use_subs = True
    try:
        cm = lw.compile(init_states)
    except CompileBudget:
        # inline expansion explodes (deeply nested by-name definitions, e.g. serializableSnapshotIsolation.tla):
        # compile the module-level operators as CALL/RET subroutines instead
        lw = Lowering(model, seq_cap=seq_cap, type_hint=type_hint)
        lw.use_subs = True
        cm = lw.compile(init_states)
        cm.warnings.append("operators compiled as subroutines (inline expansion exceeded its budget)")
    cm.module_name = model.module_name
    return cm


def pack_words(cm: CompiledModel, frame_words) -> np.ndarray:
    """Python mirror of tlag_pack (csrc/tlag_vm.h) for host-side encoding of initial states."""
    out = [0] * cm.W
    bitpos = 0
    for off, width, bias in cm.layout.tolist():
        v = (int(frame_words[off]) - bias) & 0xFFFFFFFF
        if width < 32 and (v >> width) != 0:
            raise CompileError(f"value {frame_words[off]} does not fit a {width}-bit slot")
        wi, sh = bitpos >> 5, bitpos & 31
        out[wi] |= (v << sh) & 0xFFFFFFFF
        if sh + width > 32:
            out[wi + 1] |= v >> (32 - sh)
        bitpos += width
    return np.array(out, dtype=np.uint32)


def unpack_words(cm: CompiledModel, words) -> list:
    st = [0] * cm.state_words_unpacked
    bitpos = 0
    w = [int(x) for x in words]
    for off, width, bias in cm.layout.tolist():
        wi, sh = bitpos >> 5, bitpos & 31
        v = w[wi] >> sh
        if sh + width > 32:
            v |= (w[wi + 1] << (32 - sh)) & 0xFFFFFFFF
        if width < 32:
            v &= (1 << width) - 1
        v &= 0xFFFFFFFF
        v = v + bias
        if width == 32:
            v = ((v + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)
        st[off] = v
        bitpos += width
    return st


def _frame_of(cm: CompiledModel, st) -> list:
    frame = [0] * cm.state_words_unpacked
    for v in cm.vars:
        r = cm.codec.rep(cm.var_types[v], st[v])
        o = cm.var_off[v]
        frame[o:o + len(r)] = r
    return frame


def encode_states(cm: CompiledModel, states) -> np.ndarray:
    """Pack host states.  Under SYMMETRY every state is first replaced by the representative the device
    uses: the image under the group whose unpacked frame is word-wise (signed) lexicographically least."""
    from .front.values import permute_value
    group = getattr(cm, "group", None) or []
    out = np.zeros((len(states), cm.W), dtype=np.uint32)
    for i, st in enumerate(states):
        frame = _frame_of(cm, st)
        for perm in group:
            cand = _frame_of(cm, {v: permute_value(st[v], perm) for v in cm.vars})
            if cand < frame:
                frame = cand
        out[i] = pack_words(cm, frame)
    return out


def decode_state(cm: CompiledModel, words) -> dict:
    frame = unpack_words(cm, words)
    return {v: cm.codec.unrep(cm.var_types[v], frame, cm.var_off[v]) for v in cm.vars}


_VERDICTS = {0: OK, 1: INVARIANT, 2: ASSERT, 3: DEADLOCK, 4: EVAL_ERROR}


def result_from_engine(cm: CompiledModel, res: dict, trace=None) -> CheckResult:
    r = CheckResult()
    r.verdict = _VERDICTS[res["verdict"]]
    r.generated = res["generated"]
    r.distinct = res["distinct"]
    r.queue = res.get("queue_left", 0)
    r.depth = res["depth"]
    r.init_states = res.get("init_states", 0)
    if r.verdict == INVARIANT:
        r.invariant = cm.invariants[res["detail"]]
    elif r.verdict == ASSERT:
        r.error_text = cm.asserts[res["detail"]][0]
        if r.error_text.startswith("\x00property:"):
            r.verdict = PROPERTY
            r.invariant = r.error_text.split(":", 1)[1]
    elif r.verdict == EVAL_ERROR:
        r.error_text = f"{TRAP_NAMES.get(res['detail'], 'trap ' + str(res['detail']))} (source line {res['detail2']})"
    if trace is not None:
        states, acts = trace
        for w, a in zip(states, acts):
            act = None
            if a >= 0:
                nm, loc, mod = cm.actions[a]
                act = ("fixed", nm, loc, mod)
            r.trace.append((decode_state(cm, w), act))
    return r
