"""Fixed-width state-vector bytecode (the ISA shared by the CUDA engine and the CPU
bytecode oracle; mirrored in csrc/tlag_vm.h -- keep the two tables in sync).

One instruction = 64 bits:  op[0:8] a[8:22] b[22:36] c[36:50] d[50:64].
Operands a..d are 14-bit frame word indices unless noted; imm28(b,c) = bits 22..49,
imm28(c,d) = bits 36..63 (signed).  The frame is a per-thread array of int32 words:
[ current state (unpacked) | primed state (unpacked) | temporaries ].
"""
from __future__ import annotations

import numpy as np

OPS = [
    "HALT", "ADD", "SUB", "MUL", "LT", "LE", "EQ", "NE", "AND",
    "OR", "ADDI", "MULI", "EQI", "NEI", "LTI", "LEI", "GTI", "GEI",
    "SHRI", "ANDI", "JEQ", "JNE", "JLT", "JGE", "JEQI", "JNEI", "JLTI",
    "JGEI", "JZ", "JNZ", "JNEG", "JGEZ", "JBT", "JBF", "JBTI", "JBFI",
    "JMP", "LI", "LIW", "MOV", "MOVN", "ZERO", "LDC", "DIV", "MOD",
    "NEG", "EQN", "NOT", "LDX", "STX", "TBL", "TBLT", "BSET", "BCLR",
    "BTEST", "BOR", "BAND", "BANDN", "BISZ", "BSUB", "BCNT", "BNEXT", "BFILL",
    "BSETI", "BTESTI", "UCLAMP", "TRAP", "EMIT", "GEN", "ASSERTF", "INVF", "MADI", "BANDC", "LEXLT",
    "SFIND", "SINS", "EMITD", "CALL", "RET",
]
OP = {n: i for i, n in enumerate(OPS)}

# operand formats: which fields are registers(r) / imm28 in (b,c) 'I' / imm28 in (c,d) 'J' / small imm 'n'
FMT = {
    "HALT": "", "LI": "rI", "LIW": "rI", "MOV": "rr", "MOVN": "rrn", "ZERO": "rn", "LDC": "rIn",
    "ADD": "rrr", "SUB": "rrr", "MUL": "rrr", "DIV": "rrr", "MOD": "rrr", "NEG": "rr",
    "LT": "rrr", "LE": "rrr", "EQ": "rrr", "NE": "rrr", "EQN": "rrrn", "NOT": "rr", "AND": "rrr", "OR": "rrr",
    "LDX": "rrrn", "STX": "rrrn", "TBL": "rIr",
    "BSET": "rr", "BCLR": "rr", "BTEST": "rrr", "BOR": "rrrn", "BAND": "rrrn", "BANDN": "rrrn",
    "BISZ": "rrn", "BSUB": "rrrn", "BCNT": "rrn", "BNEXT": "rrrn", "BFILL": "rn",
    "JMP": "_I", "JZ": "rI", "JNZ": "rI", "JNEG": "rI",
    "TRAP": "nI", "EMIT": "_I", "GEN": "", "ASSERTF": "_I", "INVF": "_I",
    "ADDI": "rrJ", "MULI": "rrJ", "EQI": "rrJ", "NEI": "rrJ", "LTI": "rrJ", "LEI": "rrJ", "GTI": "rrJ",
    "GEI": "rrJ", "UCLAMP": "rI", "BSETI": "rI", "BTESTI": "rrJ", "SHRI": "rrJ", "ANDI": "rrJ", "TBLT": "rIr",
    "JEQ": "rrJ", "JNE": "rrJ", "JLT": "rrJ", "JGE": "rrJ",
    "JEQI": "rkJ", "JNEI": "rkJ", "JLTI": "rkJ", "JGEI": "rkJ",
    "JBT": "rrJ", "JBF": "rrJ", "JBTI": "rnJ", "JBFI": "rnJ", "JGEZ": "rI", "MADI": "rkr", "BANDC": "rrJ", "LEXLT": "rrrn",
    "SFIND": "rrrn", "SINS": "rrrn", "EMITD": "nI", "CALL": "rI", "RET": "r",
}

INVERSE = {"JZ": "JNZ", "JEQ": "JNE", "JLT": "JGE", "JEQI": "JNEI", "JLTI": "JGEI", "JBT": "JBF", "JBTI": "JBFI",
           "JNEG": "JGEZ"}
INVERSE.update({v: k for k, v in list(INVERSE.items())})
COND_JUMPS = set(INVERSE)
IMM14_MIN, IMM14_MAX = -(1 << 13), (1 << 13) - 1

TRAP_EVAL, TRAP_OVERFLOW, TRAP_CASE, TRAP_CHOOSE, TRAP_ASSIGN = 1, 2, 3, 4, 5
TRAP_NAMES = {1: "evaluation error (function applied outside its domain / missing record field)",
              2: "value outside its inferred fixed-width type (capacity overflow)",
              3: "CASE: no arm is true", 4: "CHOOSE: no element satisfies the predicate",
              5: "internal: unassigned variable"}

MAXREG = (1 << 14) - 1
IMM28_MIN, IMM28_MAX = -(1 << 27), (1 << 27) - 1


class Label:
    __slots__ = ("name", "pos")
    _n = 0

    def __init__(self, hint="L"):
        Label._n += 1
        self.name = f"{hint}{Label._n}"
        self.pos = None

    def __repr__(self):
        return self.name


class AsmError(Exception):
    pass


class Asm:
    def __init__(self):
        self.code = []       # list of tuples (op, a, b, c, d) with Label operands allowed / ('label', L)
        self.cpool = []      # int32 constants
        self._cp_cache = {}
        self.cur_line = 0    # source line of the construct being lowered (debug / profiling info)
        self.lines = {}      # id(instruction tuple) -> line

    def emit(self, op, *args):
        ins = (op,) + tuple(args)
        self.code.append(ins)
        self.lines[id(ins)] = self.cur_line

    def label(self, L: Label):
        self.code.append(("label", L))

    def const_table(self, words):
        """Intern a table of int32 words in the constant pool; returns its base index."""
        key = tuple(int(w) for w in words)
        base = self._cp_cache.get(key)
        if base is None:
            base = len(self.cpool)
            self.cpool.extend(key)
            self._cp_cache[key] = base
        return base

    def capture(self):
        return _Capture(self)

    def splice(self, buf):
        self.code.extend(buf)

    def assemble(self, entry_points: dict):
        """Resolve labels -> (uint64 code array, int32 cpool array, {name: pc})."""
        self.peephole()
        pos = 0
        for ins in self.code:
            if ins[0] == "label":
                ins[1].pos = pos
            else:
                pos += 1
        out = np.zeros(pos, dtype=np.uint64)
        self.line_table = np.zeros(pos, dtype=np.int32)
        i = 0
        for ins in self.code:
            if ins[0] == "label":
                continue
            out[i] = self._encode(ins)
            self.line_table[i] = self.lines.get(id(ins), 0)
            i += 1
        cp = np.array(self.cpool if self.cpool else [0], dtype=np.int64)
        cp = ((cp + (1 << 31)) % (1 << 32) - (1 << 31)).astype(np.int32)
        entries = {k: (v.pos if isinstance(v, Label) else v) for k, v in entry_points.items()}
        return out, cp, entries

    def peephole(self):
        """Branch clean-up on the symbolic code: drop `JMP L` when L is the next instruction, turn
        `Jcc L1; JMP L2; L1:` into `J!cc L2`, thread jumps through `L: JMP L'`."""
        def jump_target(ins):
            if ins[0] == "JMP":
                return ins[1]
            if ins[0] in COND_JUMPS:
                return ins[-1]
            return None

        self.fuse_compare_branch()
        for _ in range(4):
            code = self.code
            # label -> first real instruction after it (for threading)
            nxt_real = {}
            pending = []
            for ins in code:
                if ins[0] == "label":
                    pending.append(ins[1])
                else:
                    for L in pending:
                        nxt_real[id(L)] = ins
                    pending = []
            out = []
            changed = False
            i, n = 0, len(code)
            while i < n:
                ins = code[i]
                if ins[0] == "label":
                    out.append(ins)
                    i += 1
                    continue
                tgt = jump_target(ins)
                if tgt is not None:
                    # thread through unconditional jumps
                    hops = 0
                    t2 = nxt_real.get(id(tgt))
                    while t2 is not None and t2[0] == "JMP" and t2[1] is not tgt and hops < 8:
                        tgt = t2[1]
                        t2 = nxt_real.get(id(tgt))
                        hops += 1
                    if hops:
                        new = ins[:-1] + (tgt,) if ins[0] != "JMP" else ("JMP", tgt)
                        self.lines[id(new)] = self.lines.get(id(ins), 0)
                        ins = new
                        changed = True
                    # labels immediately following
                    j = i + 1
                    following = []
                    while j < n and code[j][0] == "label":
                        following.append(code[j][1])
                        j += 1
                    if ins[0] == "JMP" and any(L is tgt for L in following):
                        changed = True
                        i += 1
                        continue
                    if ins[0] in COND_JUMPS and i + 1 < n and code[i + 1][0] == "JMP":
                        k = i + 2
                        foll2 = []
                        while k < n and code[k][0] == "label":
                            foll2.append(code[k][1])
                            k += 1
                        if any(L is tgt for L in foll2):
                            new = (INVERSE[ins[0]],) + ins[1:-1] + (code[i + 1][1],)
                            self.lines[id(new)] = self.lines.get(id(ins), 0)
                            out.append(new)
                            changed = True
                            i += 2
                            continue
                out.append(ins)
                i += 1
            self.code = out
            if not changed:
                break

    def fuse_compare_branch(self):
        """`CMP t, x, y ; JNZ/JZ t, L`  ->  one fused compare-and-branch.  Every compare the lowering
        emits writes a freshly allocated temporary that is consumed only by the jump that follows."""
        out = []
        code = self.code
        i, n = 0, len(code)
        fits = lambda v: isinstance(v, int) and IMM14_MIN <= v <= IMM14_MAX
        while i < n:
            a = code[i]
            b = code[i + 1] if i + 1 < n else None
            new = None
            if b is not None and b[0] in ("JNZ", "JZ") and a[0] != "label" and len(a) >= 2 and b[1] == a[1]:
                t_ = b[0] == "JNZ"
                L = b[2]
                op = a[0]
                if op == "EQ":
                    new = ("JEQ" if t_ else "JNE", a[2], a[3], L)
                elif op == "NE":
                    new = ("JNE" if t_ else "JEQ", a[2], a[3], L)
                elif op == "LT":
                    new = ("JLT" if t_ else "JGE", a[2], a[3], L)
                elif op == "LE":
                    new = ("JGE" if t_ else "JLT", a[3], a[2], L)
                elif op == "EQI" and fits(a[3]):
                    new = ("JEQI" if t_ else "JNEI", a[2], a[3], L)
                elif op == "NEI" and fits(a[3]):
                    new = ("JNEI" if t_ else "JEQI", a[2], a[3], L)
                elif op == "LTI" and fits(a[3]):
                    new = ("JLTI" if t_ else "JGEI", a[2], a[3], L)
                elif op == "GEI" and fits(a[3]):
                    new = ("JGEI" if t_ else "JLTI", a[2], a[3], L)
                elif op == "LEI" and fits(a[3] + 1):
                    new = ("JLTI" if t_ else "JGEI", a[2], a[3] + 1, L)
                elif op == "GTI" and fits(a[3] + 1):
                    new = ("JGEI" if t_ else "JLTI", a[2], a[3] + 1, L)
                elif op == "BTEST":
                    new = ("JBT" if t_ else "JBF", a[2], a[3], L)
                elif op == "BTESTI" and isinstance(a[3], int) and 0 <= a[3] <= MAXREG:
                    new = ("JBTI" if t_ else "JBFI", a[2], a[3], L)
            if new is not None:
                self.lines[id(new)] = self.lines.get(id(a), 0)
                out.append(new)
                i += 2
            else:
                out.append(a)
                i += 1
        self.code = out

    def _encode(self, ins):
        name = ins[0]
        fmt = FMT[name]
        args = list(ins[1:])
        a = b = c = d = 0
        fields = []
        ai = 0
        slot = 0  # 0:a 1:b 2:c 3:d

        def val(x):
            if isinstance(x, Label):
                if x.pos is None:
                    raise AsmError(f"unresolved label {x}")
                return x.pos
            return int(x)

        regs = [0, 0, 0, 0]
        for ch in fmt:
            if ch == "_":
                slot += 1
                continue
            v = val(args[ai])
            ai += 1
            if ch in "rn":
                if not (0 <= v <= MAXREG):
                    raise AsmError(f"{name}: operand {v} out of 14-bit range")
                regs[slot] = v
                slot += 1
            elif ch == "k":
                if not (-(1 << 13) <= v < (1 << 13)):
                    raise AsmError(f"{name}: immediate {v} out of signed 14-bit range")
                regs[slot] = v & 0x3FFF
                slot += 1
            elif ch == "I":
                if not (IMM28_MIN <= v <= IMM28_MAX):
                    raise AsmError(f"{name}: immediate {v} out of 28-bit range")
                v &= (1 << 28) - 1
                regs[1] = v & 0x3FFF
                regs[2] = (v >> 14) & 0x3FFF
                slot = 3
            elif ch == "J":
                if not (IMM28_MIN <= v <= IMM28_MAX):
                    raise AsmError(f"{name}: immediate {v} out of 28-bit range")
                v &= (1 << 28) - 1
                regs[2] = v & 0x3FFF
                regs[3] = (v >> 14) & 0x3FFF
                slot = 4
        if ai != len(args):
            raise AsmError(f"{name}: expected {ai} operands, got {len(args)}")
        a, b, c, d = regs
        return np.uint64(OP[name] | (a << 8) | (b << 22) | (c << 36) | (d << 50))


class _Capture:
    def __init__(self, asm):
        self.asm = asm

    def __enter__(self):
        self.saved = self.asm.code
        self.asm.code = []
        return self

    def __exit__(self, *exc):
        self.buf = self.asm.code
        self.asm.code = self.saved
        return False


def disasm(code, entries=None):
    inv = {}
    if entries:
        for k, v in entries.items():
            inv.setdefault(v, []).append(k)
    lines = []
    for pc, w in enumerate(code):
        w = int(w)
        op = w & 0xFF
        a = (w >> 8) & 0x3FFF
        b = (w >> 22) & 0x3FFF
        c = (w >> 36) & 0x3FFF
        d = (w >> 50) & 0x3FFF
        name = OPS[op]
        fmt = FMT[name]
        regs = [a, b, c, d]
        outs = []
        slot = 0
        for ch in fmt:
            if ch == "_":
                slot += 1
            elif ch in "rn":
                outs.append(("r" if ch == "r" else "#") + str(regs[slot]))
                slot += 1
            elif ch == "k":
                v = regs[slot]
                outs.append("$" + str(v - (1 << 14) if v >= (1 << 13) else v))
                slot += 1
            elif ch == "I":
                v = b | (c << 14)
                if v >= (1 << 27):
                    v -= 1 << 28
                outs.append(f"${v}")
                slot = 3
            elif ch == "J":
                v = c | (d << 14)
                if v >= (1 << 27):
                    v -= 1 << 28
                outs.append(f"${v}")
                slot = 4
        for nm in inv.get(pc, []):
            lines.append(f"{nm}:")
        lines.append(f"{pc:5d}  {name:8s} " + ", ".join(outs))
    return "\n".join(lines)
