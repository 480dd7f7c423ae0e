"""Sliced native build (tla_rust_b200/compile/sliced.py): one C / CUDA function per invariant and per disjunct of Next.

CPU side (no GPU): the generated code, compiled by gcc into the CPU bytecode engine in place of its interpreter
(oracle/tlag_cpu.c: TLAG_SLICED_INC -- every slice starts from the packed state and a poisoned frame, as on the
device), must reproduce the recorded oracle results of the fixtures bit for bit, in both frame forms.  The templates of
sliced.py restate the ISA independently of csrc/tlag_vm_exec.inc, so this also cross-checks the two implementations of
every opcode the fixtures use.  GPU side (-m gpu): the same slices as CUDA kernels behind the C ABI."""
import ctypes as C
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from oracle import cpu_engine
from tla_rust_b200.compiled import load_compiled
from tla_rust_b200.compile.sliced import Emitter, Plan, SliceError, emit_sliced, model_key

KEYS = ("verdict", "generated", "distinct", "depth", "fp_xor", "fp_sum", "levels")


def _cpu_sliced_lib(tmp_path, cm, scalar):
    inc = tmp_path / f"{model_key(cm)}{'_s' if scalar else ''}.inc"
    inc.write_text(emit_sliced(cm, scalar=scalar))
    so = tmp_path / (inc.stem + ".so")
    subprocess.check_call(["gcc", "-O1", "-std=gnu11", "-fPIC", "-shared", "-pthread", "-Wno-unused-label",
                           "-Wno-unused-function", f'-DTLAG_SLICED_INC="{inc}"', "-o", str(so),
                           os.path.join(ROOT, "oracle", "tlag_cpu.c")])
    L = C.CDLL(str(so))
    L.tlagcpu_run.restype = C.c_int
    L.tlagcpu_run2.restype = C.c_int
    return L


def _run_with(L, cm, init, info, **kw):
    cpu_engine.lib()
    saved, cpu_engine._LIB = cpu_engine._LIB, L
    try:
        return cpu_engine.run(cm, init, n_threads=2, deadlock=info["deadlock"], **kw)
    finally:
        cpu_engine._LIB = saved


# all verdict kinds (ok, Assert failure, invariant violation, deadlock), PlusCal, sequences, bitset-heavy Paxos,
# refinement PROPERTY + SYMMETRY (canonicalisation subroutine), records / CHOOSE / containers
@pytest.mark.parametrize("name", ["atomic_add", "pcal_intro", "pcal_intro_readme_buggy", "demo_race", "demo_lock",
                                  "MCInnerFIFO", "MCAlternatingBit", "MCPaxos", "MCPaxos3", "MCPaxos3_sym", "MCVoting",
                                  "MCVoting_deadlock", "Containers", "HourClock", "AsynchInterface", "MCPaxos3_b2"])
def test_sliced_code_reproduces_the_fixture_on_the_cpu_engine(tmp_path, name):
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    forms = [False]
    try:
        Emitter(cm, scalar=True)
        forms.append(True)
    except SliceError:
        pass
    for scalar in forms:
        r = _run_with(_cpu_sliced_lib(tmp_path, cm, scalar), cm, init, info)
        for k in KEYS:
            assert r[k] == exp["o2"][k], (name, "scalar" if scalar else "array", k)


def test_sliced_code_at_eight_million_states(tmp_path):
    """Beyond the AST oracle's reach (~10^5 states) the interpreter-side check is VM against VM; the emitted C is a second
    implementation of every opcode and of pack / unpack: MCPaxos3 with ballots 0..3 (8,220,065 states) through the scalar
    form inside the CPU engine must land on the digest ORACLE O2's interpreter recorded."""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "MCPaxos3_b3.tlagz"))
    L = _cpu_sliced_lib(tmp_path, cm, True)
    cpu_engine.lib()
    saved, cpu_engine._LIB = cpu_engine._LIB, L
    try:
        r = cpu_engine.run(cm, init, n_threads=os.cpu_count() or 2, deadlock=info["deadlock"], max_states=1 << 24)
    finally:
        cpu_engine._LIB = saved
    for k in KEYS:
        assert r[k] == exp["o2"][k], k


def test_subroutines_and_sparse_containers_in_sliced_code(tmp_path):
    """CALL -> a C function per subroutine, RET -> return; SFIND / SINS; raft and SSI at their smallest bounds."""
    for name in ("MCraft", "MCssi"):
        cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
        r = _run_with(_cpu_sliced_lib(tmp_path, cm, False), cm, init, info)
        for k in KEYS:
            assert r[k] == exp["o2"][k], (name, k)


def test_slices_partition_the_programs_and_are_closed_under_their_jumps():
    for name in ("MCPaxos3_b4", "MCraft_t4l3", "MCssi_2x2", "pcal_intro"):
        cm, _, _, _ = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
        for min_slice in (0, 256):
            pl = Plan(cm, min_slice)
            for prog, ((p0, p1), segs) in pl.progs.items():
                assert p0 == cm.entries[prog] and segs[0][0] == p1
                for (s, e), (s2, _) in zip(segs, segs[1:]):
                    assert s < e == s2
                assert pl.ins[segs[-1][1]].op == "HALT"
                for s, e in segs:
                    for i in pl.ins[s:e]:
                        t = i.target()
                        assert t is None or i.op == "CALL" or s <= t <= e, (name, prog, i.k, t)
            # Paxos: one slice per instance of Phase1a / Phase2a / Phase1b / Phase2b, grouped into kernels of >= 256
            # instructions; raft: per server / pair / message rule
            if name == "MCPaxos3_b4":
                assert (len(pl.progs["next"][1]), len(pl.progs["inv"][1])) == ((21, 4) if min_slice == 0 else (7, 2))
            if name == "MCraft_t4l3" and min_slice == 0:
                assert len(pl.progs["next"][1]) >= 30


def test_scalar_form_keeps_only_dynamically_indexed_regions_in_memory():
    cm, _, _, _ = load_compiled(os.path.join(GOLDEN, "MCPaxos3_b4.tlagz"))
    src = emit_sliced(cm, scalar=True)
    assert "TLAG_SL_SCALAR 1" in src and "f[" not in src.split("#define TLAG_SL_NEXT_LIST")[1]
    # the packed state of Paxos (the msgs bitset included) is never indexed dynamically: it lives in C locals
    e = Emitter(cm, scalar=True)
    (pro, segs) = e.plan.progs["next"]
    e._find_dyn([pro, segs[0]])
    assert not e.dyn_index and "int32_t m[" not in src       # regions of <= 16 words: select chains, no memory frame
    with pytest.raises(SliceError):
        Emitter(load_compiled(os.path.join(GOLDEN, "MCssi.tlagz"))[0], scalar=True)      # subroutines -> array form


def test_sliced_engine_library_cross_compiles_and_keeps_the_c_abi():
    from tla_rust_b200 import engine
    cm, _, _, _ = load_compiled(os.path.join(GOLDEN, "MCPaxos3.tlagz"))
    so = engine.build_sliced_library(cm)
    L = engine._bind(so)                       # raises if an include/tlag.h symbol is missing
    assert b"sliced" in L.tlag_version()
    other, _, _, _ = load_compiled(os.path.join(GOLDEN, "pcal_intro.tlagz"))
    saved = dict(engine._NATIVE_LIBS)
    try:
        engine._NATIVE_LIBS[engine.sliced_library_path(other)] = L
        with pytest.raises((engine.EngineError, engine.EngineUnavailable)) as ei:
            engine.Engine(other, native="sliced")
    finally:
        engine._NATIVE_LIBS.clear()
        engine._NATIVE_LIBS.update(saved)
    assert "another model" in str(ei.value)


# ---- device ------------------------------------------------------------------------------------------------------
SLICED_GPU = ["atomic_add", "pcal_intro", "pcal_intro_readme_buggy", "demo_race", "MCVoting_deadlock", "MCPaxos3",
              "MCPaxos3_sym", "MCInnerFIFO", "Containers", "MCraft", "MCssi", "MCPaxos3_b2", "MCPaxos3_b3"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", SLICED_GPU)
def test_sliced_kernels_match_the_oracle_on_the_device(name):
    import numpy as np
    from tla_rust_b200.engine import Engine
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    e = Engine(cm, deadlock=info["deadlock"], native="sliced")
    assert b"sliced" in e.L.tlag_version()
    e.seed(init)
    levels = [int(len(np.unique(init, axis=0)))]
    while True:
        ws = e.step()
        if ws["expanded"]:
            levels.append(int(ws["discovered"]))
        if ws["verdict"] != 5:
            break
    r = e.result()
    o2 = exp["o2"]
    assert r["verdict"] == o2["verdict"], (r, o2)
    assert (r["generated"], r["distinct"], r["depth"]) == (o2["generated"], o2["distinct"], o2["depth"])
    assert levels == o2["levels"]
    assert e.digest() == (o2["fp_xor"], o2["fp_sum"])
    if r["verdict"] in (1, 3):                 # invariant violation / deadlock: same (smallest-index) state reported
        st = e.read_states(r["state_idx"], 1)
        assert cpu_engine.digest(st, cm.W)[0] != 0
    # several kernels per level (one per group of slices), not one
    assert e.launches() >= 2 * (len(levels) - 1)
    e.close()
