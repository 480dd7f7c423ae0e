"""Model-specialised native code (tla_rust_b200/compile/native.py): the C generated from a model's bytecode, compiled
with gcc into the CPU bytecode engine in place of the interpreter (oracle/tlag_cpu.c: TLAG_NATIVE_INC), must reproduce
the recorded oracle results of the fixtures bit for bit; the CUDA engine must cross-compile with the same generated
code (nvcc, sm_100a) and keep its C ABI.  The device runs of the native build are in tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from oracle import cpu_engine
from tla_rust_b200.compiled import load_compiled
from tla_rust_b200.compile.native import emit_c, model_key


def _cpu_native_lib(tmp_path, cm, generic=False, sched="warp"):
    """sched="warp": the block form the CUDA build uses by default (one lane runs block after block here);
    "lane": the run-to-next-event form."""
    inc = tmp_path / f"{model_key(cm)}{'_g' if generic else ''}.inc"
    inc.write_text(emit_c(cm, generic=generic))
    so = tmp_path / (inc.stem + "_" + sched + ".so")
    subprocess.check_call(["gcc", "-O1", "-std=gnu11", "-fPIC", "-shared", "-pthread", "-Wno-unused-label"]
                          + (["-DTLAG_NATIVE_SCHED_WARP"] if sched == "warp" else [])
                          + [f'-DTLAG_NATIVE_INC="{inc}"', "-o", str(so), os.path.join(ROOT, "oracle", "tlag_cpu.c")])
    L = C.CDLL(str(so))
    L.tlagcpu_run.restype = C.c_int
    L.tlagcpu_probe_batch.restype = C.c_double
    L.tlagcpu_fingerprint.restype = C.c_uint64
    return L


def _run_with(L, cm, init, info):
    cpu_engine.lib()
    saved, cpu_engine._LIB = cpu_engine._LIB, L
    try:
        return cpu_engine.run(cm, init, n_threads=2, deadlock=info["deadlock"])
    finally:
        cpu_engine._LIB = saved


# verdicts: ok, Assert failure (README transcript model), invariant violation, deadlock-free PlusCal, sequences,
# bitset-heavy Paxos, containers/records/CHOOSE
@pytest.mark.parametrize("name", ["atomic_add", "pcal_intro", "pcal_intro_readme_buggy", "demo_race", "demo_lock",
                                  "MCInnerFIFO", "MCAlternatingBit", "MCPaxos3", "Containers", "HourClock"])
def test_native_code_reproduces_the_fixture_on_the_cpu_engine(tmp_path, name):
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    for sched in ("warp", "lane"):
        r = _run_with(_cpu_native_lib(tmp_path, cm, sched=sched), cm, init, info)
        for k in ("verdict", "generated", "distinct", "depth", "fp_xor", "fp_sum", "levels"):
            assert r[k] == exp["o2"][k], (name, sched, k)


def test_direct_templates_agree_with_the_inlined_executor_form(tmp_path):
    """generic=True emits every instruction as a call of the single ISA definition with a literal instruction word."""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "MCAlternatingBit.tlagz"))
    src = emit_c(cm, generic=True)
    assert "TLAG_NATIVE_X(0x" in src and "TLAG_NATIVE_X(0x" not in emit_c(cm)
    r = _run_with(_cpu_native_lib(tmp_path, cm, generic=True), cm, init, info)
    for k in ("verdict", "generated", "distinct", "depth", "fp_xor", "fp_sum"):
        assert r[k] == exp["o2"][k], k


def test_subroutines_sparse_containers_and_symmetry_ops_in_native_code(tmp_path):
    """CALL/RET (return through the resume switch), SFIND/SINS, LEXLT: raft and SSI at their smallest bounds."""
    for name, sched in (("MCraft", "warp"), ("MCssi", "warp"), ("MCPaxos3_sym", "warp"), ("MCraft", "lane")):
        cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
        r = _run_with(_cpu_native_lib(tmp_path, cm, sched=sched), cm, init, info)
        for k in ("verdict", "generated", "distinct", "depth", "fp_xor", "fp_sum"):
            assert r[k] == exp["o2"][k], (name, sched, k)


def test_native_engine_library_cross_compiles_and_keeps_the_c_abi():
    from tla_rust_b200 import engine
    cm, _, _, _ = load_compiled(os.path.join(GOLDEN, "MCPaxos3.tlagz"))
    so = engine.build_native_library(cm)
    L = engine._bind(so)                       # raises if an include/tlag.h symbol is missing
    assert b"native" in L.tlag_version()
    # the library refuses any other model's program (checked before any CUDA call)
    other, _, _, _ = load_compiled(os.path.join(GOLDEN, "pcal_intro.tlagz"))
    with pytest.raises((engine.EngineError, engine.EngineUnavailable)) as ei:
        saved = dict(engine._NATIVE_LIBS)
        try:
            engine._NATIVE_LIBS[engine.native_library_path(other)] = L
            engine.Engine(other, native=True)
        finally:
            engine._NATIVE_LIBS.clear()
            engine._NATIVE_LIBS.update(saved)
    assert "another model" in str(ei.value)
