"""ISA-level tests of the extension ops (csrc/tlag_vm_exec.inc): hand-assembled programs run by the CPU bytecode
engine, which executes the same tlag_vm_exec as the CUDA interpreter.  Each program is a Next action over a small
unpacked state; the successor it EMITs is read back and compared with a Python model of the op."""
import numpy as np
import pytest

from tla_rust_b200.compile.bytecode import Asm, Label
from tla_rust_b200.compile.lower import CompiledModel
from oracle import cpu_engine


def _model(n_words, build, frame=256):
    """State = n_words unpacked 32-bit slots (current at 0.., primed at n_words..); `build(asm, cur, nxt, tmp)`
    emits the body of Next.  Returns the list of distinct states reached from the all-zero state."""
    a = Asm()
    inv, nxt = Label("inv"), Label("next")
    a.label(inv)
    a.emit("HALT")
    a.label(nxt)
    build(a, 0, n_words, 2 * n_words)
    a.emit("HALT")
    code, cpool, ent = a.assemble({"inv": inv, "next": nxt})
    cm = CompiledModel()
    cm.code, cm.cpool, cm.entries = code, cpool, ent
    cm.W, cm.frame_words, cm.state_words_unpacked = n_words, frame, n_words
    cm.layout = np.array([[i, 32, 0] for i in range(n_words)], dtype=np.int32)
    cm.invariants, cm.actions = [], [("Next", (0, 0, 0, 0), "T")]
    init = np.zeros((1, n_words), dtype=np.uint32)
    r = cpu_engine.run(cm, init, deadlock=False, want_states=True, max_states=1 << 10)
    assert r["verdict"] == 0, r
    return [tuple(int(x) for x in row.astype(np.int32)) for row in r["states"]]


def test_sins_keeps_entries_sorted_and_sfind_finds_them():
    # container at primed[0..]: length + cap 4 entries of (key, value); stride 2, key width 1
    desc = (2 << 7) | 1
    n = 1 + 4 * 2 + 1                      # + one slot for the SFIND result

    def build(a, cur, nxt, tmp):
        a.emit("MOVN", nxt, cur, n)
        for k, v in ((5, 50), (3, 30), (9, 90), (5, 55)):      # the second 5 overwrites the value of key 5
            a.emit("LI", tmp, k)
            a.emit("LI", tmp + 1, v)
            a.emit("LI", tmp + 2, 4)                            # capacity in, status out
            a.emit("SINS", nxt, tmp, tmp + 2, desc)
        a.emit("LI", tmp, 9)
        a.emit("SFIND", nxt + 9, nxt, tmp, desc)                # index of key 9 -> 2
        a.emit("EMIT", 0)
    states = _model(n, build)
    assert (3, 3, 30, 5, 55, 9, 90, 0, 0, 2) in states


def test_sins_reports_a_full_container_and_sfind_misses():
    desc = (1 << 7) | 1
    n = 1 + 2 + 2

    def build(a, cur, nxt, tmp):
        a.emit("MOVN", nxt, cur, n)
        for k in (7, 4, 6):                                     # capacity 2: the third insert must fail
            a.emit("LI", tmp, k)
            a.emit("LI", tmp + 1, 2)
            a.emit("SINS", nxt, tmp, tmp + 1, desc)
        a.emit("MOV", nxt + 3, tmp + 1)                         # status of the last insert: 0
        a.emit("LI", tmp, 5)
        a.emit("SFIND", nxt + 4, nxt, tmp, desc)                # 5 is absent: -1
        a.emit("EMIT", 0)
    states = _model(n, build)
    assert (2, 4, 7, 0, -1) in states


def test_call_ret_and_lexlt():
    def build(a, cur, nxt, tmp):
        sub, over = Label("sub"), Label("over")
        a.emit("MOVN", nxt, cur, 3)
        a.emit("LI", tmp + 1, 20)
        a.emit("CALL", tmp, sub)                                # nxt[0] = arg + 1
        a.emit("LI", tmp + 1, 40)
        a.emit("CALL", tmp, sub)                                # nxt[0] = 41 (second activation overwrites)
        a.emit("LI", tmp + 2, 1)
        a.emit("LI", tmp + 3, 2)
        a.emit("LI", tmp + 4, 1)
        a.emit("LI", tmp + 5, 3)
        a.emit("LEXLT", nxt + 1, tmp + 2, tmp + 4, 2)            # (1,2) < (1,3) -> 1
        a.emit("LEXLT", nxt + 2, tmp + 4, tmp + 2, 2)            # (1,3) < (1,2) -> 0
        a.emit("EMIT", 0)
        a.emit("JMP", over)
        a.label(sub)
        a.emit("ADDI", nxt, tmp + 1, 1)
        a.emit("RET", tmp)
        a.label(over)
    states = _model(3, build)
    assert (41, 1, 0) in states


def test_emitd_repacks_only_the_listed_slot_ranges():
    """EMITD starts from the parent's packed words: a primed slot outside the listed ranges is NOT written back,
    which is exactly the contract the compiler relies on (unlisted variables are unchanged)."""
    def build(a, cur, nxt, tmp):
        a.emit("MOVN", nxt, cur, 4)
        a.emit("LI", nxt + 1, 11)
        a.emit("LI", nxt + 2, 22)
        a.emit("LI", nxt + 3, 33)                                # not listed below: must not reach the successor
        base = a.const_table([0])                                # index 0 = "no table"
        tbl = a.const_table([1, 1, 2, 32])                       # one range: slots 1..2, bit position 32
        assert base == 0 and tbl > 0
        a.emit("EMITD", 0, tbl)
    states = _model(4, build)
    assert (0, 11, 22, 0) in states and (0, 11, 22, 33) not in states
