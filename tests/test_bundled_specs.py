"""Every bundled spec of the reference that has a .cfg (north star: "bit-exact distinct-state count and
invariant verdict ... on every bundled spec"): ORACLE O1 (AST evaluator) against the compiled model run by
ORACLE O2 (C bytecode engine).  The reference holds no expected counts for these; the pinned numbers are O1's.
Specs already covered by the committed GPU fixtures (MCPaxos, MCVoting, MCInnerFIFO, MCAlternatingBit, HourClock,
AsynchInterface) are in tests/test_compile_cpu.py / tests/test_gpu_parity.py."""
import pytest

from conftest import REF, needs_reference
from tla_rust_b200.front.spec import Model
from tla_rust_b200.checker import compile_model, encode_states
from oracle import cpu_engine
from oracle.tlc_oracle import Oracle

EX = REF + "/examples/"
SS = EX + "SpecifyingSystems/"
V = {"ok": 0, "invariant": 1, "assert": 2, "deadlock": 3}

CASES = [
    # path, seq_cap, (verdict, generated, distinct, depth)
    (EX + "Paxos/MCConsensus.tla", None, ("deadlock", 7, 4, 1)),
    (SS + "AsynchronousInterface/Channel.tla", None, ("ok", 30, 12, 2)),
    (SS + "HourClock/HourClock2.tla", None, ("ok", 24, 12, 1)),
    (SS + "Liveness/LiveHourClock.tla", None, ("ok", 24, 12, 1)),
    (SS + "TLC/ABCorrectness.tla", None, ("ok", 36, 20, 3)),
    (SS + "RealTime/MCRealTimeHourClock.tla", None, ("ok", 696, 216, 2)),          # [A]_v used as an action
    (SS + "AdvancedExamples/MCInnerSequential.tla", None, ("ok", 24368, 3528, 9)),   # Seq capacity from sampling
    (SS + "CachingMemory/MCInternalMemory.tla", None, ("ok", 21400, 4408, 10)),     # atom | record unions, sampled typing
    (SS + "Liveness/MCLiveInternalMemory.tla", None, ("ok", 21400, 4408, 10)),
    (SS + "CachingMemory/MCWriteThroughCache.tla", 2, ("ok", 28170, 5196, 18)),     # recursive local function (vmem)
]


@needs_reference
@pytest.mark.parametrize("path,seq_cap,want", CASES, ids=[c[0].split("/")[-1][:-4] for c in CASES])
def test_bundled_spec_compiled_matches_oracle(path, seq_cap, want):
    m = Model(path)
    m.check_assumes()
    init = m.initial_states()
    o1 = Oracle(m).run()
    assert (o1.verdict, o1.generated, o1.distinct, o1.depth) == want
    cm = compile_model(m, init, seq_cap=seq_cap)
    o2 = cpu_engine.run(cm, encode_states(cm, init), deadlock=m.check_deadlock)
    assert (o2["verdict"], o2["generated"], o2["distinct"], o2["depth"]) == (V[want[0]],) + want[1:]


@needs_reference
@pytest.mark.parametrize("name", ["AsynchronousInterface/PrintValues", "SimpleMath/SimpleMath"])
def test_assumption_only_modules(name):
    """No behaviour specification: TLC only evaluates the ASSUMEs (PrintValues.tla:48-54, SimpleMath.tla)."""
    m = Model(SS + name + ".tla")
    m.check_assumes()
    assert m.next_node is None and not m.init_nodes


VARIATIONS = [
    # spec, (cfg text replacement ...), seq_cap, O1's counts -- other bounds than the shipped cfg, same specs
    (SS + "FIFO/MCInnerFIFO.tla", (("qLen = 3", "qLen = 4"),), 6, ("ok", 29100, 11640, 13)),
    (SS + "TLC/MCAlternatingBit.tla", (("msgQLen = 2", "msgQLen = 3"),), 5, ("ok", 2404, 372, 11)),
    (SS + "TLC/MCAlternatingBit.tla", (("ackQLen = 2", "ackQLen = 3"),), 5, ("ok", 2212, 344, 11)),
    (SS + "CachingMemory/MCInternalMemory.tla", (("Adr = {a1, a2, a3}", "Adr = {a1, a2}"),
                                                 ("Proc = {p1, p2}", "Proc = {p1, p2, p3}")), None,
     ("ok", 153916, 23544, 13)),
]


@needs_reference
@pytest.mark.parametrize("path,subs,seq_cap,want", VARIATIONS,
                         ids=[c[0].split("/")[-1][:-4] + "-" + c[1][0][1].replace(" ", "") for c in VARIATIONS])
def test_bundled_spec_at_other_bounds(path, subs, seq_cap, want):
    cfg = open(path[:-4] + ".cfg").read()
    for a, b in subs:
        assert a in cfg
        cfg = cfg.replace(a, b)
    m = Model(path, cfg_text=cfg)
    init = m.initial_states()
    o1 = Oracle(m).run()
    assert (o1.verdict, o1.generated, o1.distinct, o1.depth) == want
    cm = compile_model(m, init, seq_cap=seq_cap)
    o2 = cpu_engine.run(cm, encode_states(cm, init), deadlock=m.check_deadlock, n_threads=2)
    assert (o2["verdict"], o2["generated"], o2["distinct"], o2["depth"]) == (V[want[0]],) + want[1:]
