"""Model-specialised native code: the bytecode of one CompiledModel turned into straight-line C / CUDA C.

Why: the wave kernel is instruction-issue bound in its interpreter loop (fetch, decode, operand-class dispatch and
the warp's min-pc election cost ~70 SASS instructions per bytecode instruction: profiles/r1_k_wave_final_b3_ncu.txt,
DESIGN.md section 4).  The program of a model is fixed for a whole run, so it can be compiled instead of interpreted.

How: every instruction at pc `K` becomes one C statement with its operand fields as literals
(`f[113] = f[97] & f[105];`, `if (f[12] == f[40]) TLAG_GOTO(731);`) -- `_direct` below restates the arithmetic of
`csrc/tlag_vm_exec.inc` op by op; branch targets are literals too.  `emit_c(generic=True)` writes the same program as
calls of the force-inlined single ISA definition (`tlag_vm_exec`) with the instruction word as a literal, which the C
compiler folds down to the same operation: slow to compile, identical by construction, and the cross-check of the
templates.  Either way the result is checked end to end by running the generated code inside the CPU bytecode engine
against the recorded digests of the fixtures (tests/test_native.py).

The generated file holds the program once, in two selectable forms (same statements, different control macros):
  * block form (-DTLAG_NATIVE_SCHED_WARP, the default of the CUDA build): `tlag_native_block` runs the lanes sitting at
    one block leader up to the next merge point (jump target / resumable pc), taken branch or event; the engine elects
    the warp's minimum pc every round, so lanes that took different paths re-join exactly as under the interpreter
    (per-lane native code would only reconverge at the next event: lanes resuming after different EMITs would run
    one after the other) while fetch / decode / operand dispatch are gone and the election is paid per block
    (~6 instructions) instead of per instruction;
  * lane form: `int tlag_native_run(cpool, f, pc_io, info, info2)` with the contract of tlag_vm_run (tlag_vm.h): one
    lane runs from *pc_io to its next event.
`csrc/tlag_engine.cu` compiled with -DTLAG_NATIVE_INC=<file> calls it per lane instead of the warp interpreter
(`tla_rust_b200/engine.py: build_native_library`)."""
from __future__ import annotations

import hashlib

import numpy as np

from .bytecode import OP

_COND_J = {"JEQ", "JNE", "JLT", "JGE", "JEQI", "JNEI", "JLTI", "JGEI", "JBT", "JBF", "JBTI", "JBFI"}   # target: immJ
_COND_I = {"JZ", "JNZ", "JNEG", "JGEZ"}                                                                # target: immI
_EVENTS = {"HALT", "TRAP", "EMIT", "EMITD", "GEN", "ASSERTF", "INVF", "DIV", "MOD", "TBLT"}
_NAME = {v: k for k, v in OP.items()}


def _imm28(v: int) -> int:
    v &= 0xFFFFFFF
    return v - (1 << 28) if v & (1 << 27) else v


def model_key(cm) -> str:
    """Identifies the generated code: program words, entry points (the constant pool is read at run time)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(cm.code, dtype=np.uint64).tobytes())
    h.update(repr(sorted(cm.entries.items())).encode())
    return h.hexdigest()[:16]


_BIN = {"ADD": "(int32_t)((uint32_t)f[{b}] + (uint32_t)f[{c}])", "SUB": "(int32_t)((uint32_t)f[{b}] - (uint32_t)f[{c}])",
        "MUL": "(int32_t)((uint32_t)f[{b}] * (uint32_t)f[{c}])", "LT": "f[{b}] < f[{c}]", "LE": "f[{b}] <= f[{c}]",
        "EQ": "f[{b}] == f[{c}]", "NE": "f[{b}] != f[{c}]", "AND": "(f[{b}] != 0) & (f[{c}] != 0)",
        "OR": "(f[{b}] != 0) | (f[{c}] != 0)"}
_BINI = {"ADDI": "(int32_t)((uint32_t)f[{b}] + (uint32_t)({J}))", "MULI": "(int32_t)((uint32_t)f[{b}] * (uint32_t)({J}))",
         "EQI": "f[{b}] == ({J})", "NEI": "f[{b}] != ({J})", "LTI": "f[{b}] < ({J})", "LEI": "f[{b}] <= ({J})",
         "GTI": "f[{b}] > ({J})", "GEI": "f[{b}] >= ({J})", "SHRI": "(int32_t)((uint32_t)f[{b}] >> (({J}) & 31))",
         "ANDI": "f[{b}] & ({J})"}
_CMP = {"JEQ": "==", "JNE": "!=", "JLT": "<", "JGE": ">=", "JEQI": "==", "JNEI": "!=", "JLTI": "<", "JGEI": ">=",
        "JZ": "==", "JNZ": "!=", "JNEG": "<", "JGEZ": ">="}
_WORDWISE = {"BOR": "f[{b} + i] | f[{c} + i]", "BAND": "f[{b} + i] & f[{c} + i]", "BANDN": "f[{b} + i] & ~f[{c} + i]"}


def _direct(op, k, a, b, c, d, I, J):
    """C statement(s) for instruction k written out directly -- the same arithmetic as csrc/tlag_vm_exec.inc, with the
    operand fields as literals (compiles ~20x faster than folding the inlined executor).  None: no template, use the
    generic form.  Jumps and events are handled by the caller."""
    if op in _BIN:
        return f"f[{a}] = {_BIN[op].format(b=b, c=c)};"
    if op in _BINI:
        return f"f[{a}] = {_BINI[op].format(b=b, J=J)};"
    if op == "LI":
        return f"f[{a}] = {I};"
    if op == "LIW":
        return f"f[{a}] = tlag_cp(cpool, {I});"
    if op == "MOV":
        return f"f[{a}] = f[{b}];"
    if op == "MOVN":
        if a <= b:
            return f"for (uint32_t i = 0; i < {c}u; ++i) f[{a} + i] = f[{b} + i];"
        return f"for (uint32_t i = {c}u; i-- > 0;) f[{a} + i] = f[{b} + i];"
    if op == "ZERO":
        return f"for (uint32_t i = 0; i < {b}u; ++i) f[{a} + i] = 0;"
    if op == "LDC":
        return f"for (uint32_t i = 0; i < {d}u; ++i) f[{a} + i] = tlag_cp(cpool, {I} + (int32_t)i);"
    if op == "NEG":
        return f"f[{a}] = -f[{b}];"
    if op == "NOT":
        return f"f[{a}] = !f[{b}];"
    if op == "EQN":
        return f"{{ int32_t e = 1; for (uint32_t i = 0; i < {d}u; ++i) e &= (f[{b} + i] == f[{c} + i]); f[{a}] = e; }}"
    if op == "LDX":
        return (f"{{ const uint32_t base = {b}u + (uint32_t)f[{c}] * {d}u; "
                f"for (uint32_t i = 0; i < {d}u; ++i) f[{a} + i] = f[base + i]; }}")
    if op == "STX":
        return (f"{{ const uint32_t base = {a}u + (uint32_t)f[{b}] * {d}u; "
                f"for (uint32_t i = 0; i < {d}u; ++i) f[base + i] = f[{c} + i]; }}")
    if op == "TBL":
        return f"f[{a}] = tlag_cp(cpool, {I} + f[{d}]);"
    if op == "BSET":
        return f"{{ const uint32_t i = (uint32_t)f[{b}]; f[{a} + (i >> 5)] |= (int32_t)(1u << (i & 31)); }}"
    if op == "BCLR":
        return f"{{ const uint32_t i = (uint32_t)f[{b}]; f[{a} + (i >> 5)] &= ~(int32_t)(1u << (i & 31)); }}"
    if op == "BTEST":
        return (f"{{ const uint32_t i = (uint32_t)f[{c}]; "
                f"f[{a}] = (int32_t)(((uint32_t)f[{b} + (i >> 5)] >> (i & 31)) & 1u); }}")
    if op in _WORDWISE:
        return f"for (uint32_t i = 0; i < {d}u; ++i) f[{a} + i] = {_WORDWISE[op].format(b=b, c=c)};"
    if op == "BISZ":
        return f"{{ int32_t z = 1; for (uint32_t i = 0; i < {c}u; ++i) z &= (f[{b} + i] == 0); f[{a}] = z; }}"
    if op == "BSUB":
        return (f"{{ int32_t z = 1; for (uint32_t i = 0; i < {d}u; ++i) z &= ((f[{b} + i] & ~f[{c} + i]) == 0); "
                f"f[{a}] = z; }}")
    if op == "BCNT":
        return (f"{{ int32_t n = 0; for (uint32_t i = 0; i < {c}u; ++i) n += tlag_popc((uint32_t)f[{b} + i]); "
                f"f[{a}] = n; }}")
    if op == "BSETI":
        i = I & 0xFFFFFFFF
        return f"f[{a + (i >> 5)}] |= (int32_t)0x{1 << (i & 31):x}u;"
    if op == "BTESTI":
        i = J & 0xFFFFFFFF
        return f"f[{a}] = (int32_t)(((uint32_t)f[{b + (i >> 5)}] >> {i & 31}) & 1u);"
    if op == "UCLAMP":
        return f"if ((uint32_t)f[{a}] >= (uint32_t)({I})) f[{a}] = -1;"
    if op == "MADI":
        k14 = b - (1 << 14) if b & (1 << 13) else b
        return f"f[{a}] = (int32_t)((uint32_t)f[{a}] * (uint32_t)({k14}) + (uint32_t)f[{c}]);"
    if op == "BANDC":
        n, base = J & 0xFF, (J & 0xFFFFFFFF) >> 8
        return f"for (uint32_t i = 0; i < {n}u; ++i) f[{a} + i] = f[{b} + i] & tlag_cp(cpool, {base} + (int32_t)i);"
    trap = f"{{ *info = 1; *info2 = 0; *pc_io = {k + 1}u; return TLAG_EV_TRAP; }}"
    if op == "DIV":
        return (f"{{ const int32_t x = f[{b}], y = f[{c}]; if (y == 0) {trap} "
                f"int32_t q = x / y; if ((x % y != 0) && ((x < 0) != (y < 0))) --q; f[{a}] = q; }}")
    if op == "MOD":
        return (f"{{ const int32_t x = f[{b}], y = f[{c}]; if (y <= 0) {trap} "
                f"int32_t r = x % y; if (r < 0) r += y; f[{a}] = r; }}")
    if op == "TBLT":
        return (f"{{ const int32_t v = tlag_cp(cpool, {I} + f[{d}]); if (v == (int32_t)0x80000000) {trap} f[{a}] = v; }}")
    if op == "BNEXT":
        return (f"{{ int32_t cur = f[{c}] + 1; int32_t res = -1; "
                f"while ((uint32_t)cur < {d}u) {{ "
                f"const uint32_t word = (uint32_t)f[{b} + ((uint32_t)cur >> 5)] >> ((uint32_t)cur & 31); "
                f"if (word) {{ const int32_t cand = cur + tlag_ffs(word) - 1; if ((uint32_t)cand < {d}u) res = cand; break; }} "
                f"cur = (int32_t)(((uint32_t)cur | 31u) + 1u); }} f[{a}] = res; }}")
    if op == "BFILL":
        return f"for (uint32_t i = 0; i < {b}u; ++i) f[{a} + (i >> 5)] |= (int32_t)(1u << (i & 31));"
    if op == "LEXLT":
        return (f"{{ int32_t r = 0; for (uint32_t i = 0; i < {d}u; ++i) {{ const int32_t x = f[{b} + i], y = f[{c} + i]; "
                f"if (x != y) {{ r = x < y; break; }} }} f[{a}] = r; }}")
    stride, keyw = d >> 7, d & 127
    if op == "SFIND":
        return (f"{{ const int32_t n = f[{b}]; int32_t r = -1; for (int32_t i = 0; i < n; ++i) {{ "
                f"const uint32_t e = {b}u + 1u + (uint32_t)i * {stride}u; uint32_t k = 0; "
                f"while (k < {keyw}u && f[e + k] == f[{c} + k]) ++k; "
                f"if (k == {keyw}u) {{ r = i; break; }} if (f[e + k] > f[{c} + k]) break; }} f[{a}] = r; }}")
    if op == "SINS":
        return (f"{{ const int32_t n = f[{a}]; const int32_t cap = f[{c}]; int32_t pos = 0; int32_t hit = 0; int full = 0; "
                f"for (; pos < n; ++pos) {{ const uint32_t e = {a}u + 1u + (uint32_t)pos * {stride}u; uint32_t k = 0; "
                f"while (k < {keyw}u && f[e + k] == f[{b} + k]) ++k; "
                f"if (k == {keyw}u) {{ hit = 1; break; }} if (f[e + k] > f[{b} + k]) break; }} "
                f"if (!hit) {{ if (n >= cap) {{ f[{c}] = 0; full = 1; }} else {{ "
                f"for (int32_t i = n; i > pos; --i) {{ const uint32_t dst = {a}u + 1u + (uint32_t)i * {stride}u; "
                f"for (uint32_t k = 0; k < {stride}u; ++k) f[dst + k] = f[dst - {stride}u + k]; }} f[{a}] = n + 1; }} }} "
                f"if (!full) {{ const uint32_t e = {a}u + 1u + (uint32_t)pos * {stride}u; "
                f"for (uint32_t k = 0; k < {stride}u; ++k) f[e + k] = f[{b} + k]; f[{c}] = 1; }} }}")
    return None


def emit_c(cm, generic: bool = False) -> str:
    """generic=True writes every instruction as a call of the inlined executor (slow to compile; the cross-check of
    the direct templates in tests/test_native.py)."""
    code = [int(x) for x in np.ascontiguousarray(cm.code, dtype=np.uint64)]
    n = len(code)
    resume = set(int(v) for v in cm.entries.values())
    targets = set()
    lines = []
    n_generic = 0
    for k, w in enumerate(code):
        op = _NAME.get(w & 0xFF, None)
        a, b, c, d = (w >> 8) & 0x3FFF, (w >> 22) & 0x3FFF, (w >> 36) & 0x3FFF, (w >> 50) & 0x3FFF
        immI = _imm28(w >> 22)
        immJ = _imm28(w >> 36)
        if op in _COND_J or op in _COND_I or op in ("JMP", "CALL"):
            t = immJ if op in _COND_J else immI
            if not 0 <= t < n:
                raise ValueError(f"branch target {t} outside the program at pc {k}")
            targets.add(t)
        if not generic:
            st = None
            if op in ("JEQ", "JNE", "JLT", "JGE"):
                st = f"if (f[{a}] {_CMP[op]} f[{b}]) TLAG_GOTO({immJ});"
            elif op in ("JEQI", "JNEI", "JLTI", "JGEI"):
                k14 = b - (1 << 14) if b & (1 << 13) else b
                st = f"if (f[{a}] {_CMP[op]} ({k14})) TLAG_GOTO({immJ});"
            elif op in _COND_I:
                st = f"if (f[{a}] {_CMP[op]} 0) TLAG_GOTO({immI});"
            elif op in ("JBT", "JBF"):
                st = (f"{{ const uint32_t i = (uint32_t)f[{b}]; if (((((uint32_t)f[{a} + (i >> 5)] >> (i & 31)) & 1u) != 0) == "
                      f"{1 if op == 'JBT' else 0}) TLAG_GOTO({immJ}); }}")
            elif op in ("JBTI", "JBFI"):
                st = (f"if (((((uint32_t)f[{a + (b >> 5)}] >> {b & 31}) & 1u) != 0) == {1 if op == 'JBTI' else 0}) "
                      f"TLAG_GOTO({immJ});")
            elif op == "JMP":
                st = f"TLAG_GOTO({immI});"
            elif op == "CALL":
                resume.add(k + 1)
                st = f"f[{a}] = {k + 1}; TLAG_GOTO({immI});"
            elif op == "RET":
                st = f"TLAG_GOTO_DYN((uint32_t)f[{a}]);"
            elif op == "HALT":
                resume.add(k)
                st = f"*pc_io = {k}u; return TLAG_EV_HALT;"
            elif op == "EMIT":
                resume.add(k + 1)
                st = f"*info = {immI}; *info2 = 0; *pc_io = {k + 1}u; return TLAG_EV_EMIT;"
            elif op == "EMITD":
                resume.add(k + 1)
                st = f"*info = {a}; *info2 = {immI}; *pc_io = {k + 1}u; return TLAG_EV_EMIT;"
            elif op == "GEN":
                resume.add(k + 1)
                st = f"*pc_io = {k + 1}u; return TLAG_EV_GEN;"
            elif op in ("ASSERTF", "INVF"):
                resume.add(k + 1)
                st = f"*info = {immI}; *pc_io = {k + 1}u; return TLAG_EV_{'ASSERT' if op == 'ASSERTF' else 'INVF'};"
            elif op == "TRAP":
                resume.add(k + 1)
                st = f"*info = {a}; *info2 = {immI}; *pc_io = {k + 1}u; return TLAG_EV_TRAP;"
            elif op is not None:
                st = _direct(op, k, a, b, c, d, immI, immJ)
                if st is not None and op in ("DIV", "MOD", "TBLT"):
                    resume.add(k + 1)
            if st is not None:
                lines.append((k, f"{{ {st} }}"))
                continue
        n_generic += 1
        call = f"{{ uint32_t pc = {k}u; const int ev = TLAG_NATIVE_X(0x{w:016x}ULL, cpool, f, &pc, info, info2);"
        if op in _COND_J or op in _COND_I:
            t = immJ if op in _COND_J else immI
            post = f" (void)ev; if (pc != {k + 1}u) TLAG_GOTO({t}); }}"
        elif op in ("JMP", "CALL"):
            if op == "CALL":
                resume.add(k + 1)
            post = f" (void)ev; (void)pc; TLAG_GOTO({immI}); }}"
        elif op == "RET":
            post = " (void)ev; TLAG_GOTO_DYN(pc); }"
        elif op in _EVENTS or op is None:
            # the executor leaves the pc to resume at in `pc` (K for HALT, K + 1 otherwise)
            resume.add(k)
            resume.add(k + 1)
            post = " if (ev >= 0) { *pc_io = pc; return ev; } }"
        else:
            post = " (void)ev; (void)pc; }"
        lines.append((k, call + post))
    resume = sorted(r for r in resume if 0 <= r < n)
    leaders = set(resume) | targets
    body = []
    for k, text in lines:
        if k in leaders:
            body.append(f"  {'TLAG_FALL(%d) ' % k if k else ''}TLAG_LABEL({k}) {text}")
        else:
            body.append(f"  {text}")
    fnv = 0xcbf29ce484222325
    for w in code:
        fnv = ((fnv ^ w) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    sig = ("(const int32_t* __restrict__ cpool, int32_t* __restrict__ f, uint32_t* pc_io,\n"
           "                                     int32_t* info, int32_t* info2) {")
    out = [
        "// generated by tla_rust_b200/compile/native.py -- do not edit",
        f"// model key {model_key(cm)}: {n} instructions ({n_generic} through the inlined executor), "
        f"{len(resume)} resumable pcs, {len(leaders)} block leaders",
        "#ifndef TLAG_NATIVE_X",
        "#error \"define TLAG_NATIVE_X (the force-inlined single-instruction executor) before including\"",
        "#endif",
        f"#define TLAG_NATIVE_CODE_LEN {n}u",
        f"#define TLAG_NATIVE_CODE_FNV 0x{fnv:016x}ULL   /* FNV-1a over the 64-bit program words */",
        "#ifdef TLAG_NATIVE_SCHED_WARP",
        "// Block form: runs the lanes that sit at *pc_io from that block leader up to the next merge point (a jump target or",
        "// a resumable pc), a taken branch, or an event; returns -1 with the successor pc in *pc_io, or the event.  The",
        "// caller elects the minimum pc of the warp each round, so lanes re-join exactly as under the interpreter.",
        "#define TLAG_LABEL(K) case K:",
        "#define TLAG_GOTO(T) do { *pc_io = T; return -1; } while (0)",
        "#define TLAG_GOTO_DYN(X) do { *pc_io = (X); return -1; } while (0)",
        "#define TLAG_FALL(K) *pc_io = K; return -1;",
        "TLAG_NATIVE_QUAL int tlag_native_block" + sig,
        "  switch (*pc_io) {",
    ]
    out += body
    out += ["    default: *info = 98; *info2 = (int32_t)*pc_io; return TLAG_EV_TRAP;",
            "  }",
            f"  *info = 97; *info2 = {n}; *pc_io = {n}u; return TLAG_EV_TRAP;", "}",
            "#else",
            "// Lane form: one lane runs from *pc_io to its next event (contract of tlag_vm_run).",
            "#define TLAG_LABEL(K) L_##K:",
            "#define TLAG_GOTO(T) goto L_##T",
            "#define TLAG_GOTO_DYN(X) do { gpc = (X); goto dispatch; } while (0)",
            "#define TLAG_FALL(K)",
            "TLAG_NATIVE_QUAL int tlag_native_run" + sig,
            "  uint32_t gpc = *pc_io;",
            "dispatch:",
            "  switch (gpc) {"]
    out += [f"    case {r}u: goto L_{r};" for r in resume]
    out += ["    default: *info = 98; *info2 = (int32_t)gpc; *pc_io = gpc; return TLAG_EV_TRAP;",
            "  }"]
    out += body
    # falling off the end of the program is a compiler bug, not a model error
    out += [f"  *info = 97; *info2 = {n}; *pc_io = {n}u; return TLAG_EV_TRAP;", "}",
            "#endif",
            "#undef TLAG_LABEL", "#undef TLAG_GOTO", "#undef TLAG_GOTO_DYN", "#undef TLAG_FALL", ""]
    return "\n".join(out)
