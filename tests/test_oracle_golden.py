"""Pins the oracle (oracle/tlc_oracle.py) against the only known-answer transcript the reference
holds for the hot path: README.md:267-321 (buggy pcal_intro) -- counts, trace, action locations --
and the verdict-only expectations (README.md:349-352, HourClock = 12 states)."""
import os
import tempfile

import pytest

from conftest import REF, ROOT, needs_reference
from tla_rust_b200.front.spec import Model
from tla_rust_b200.front.pcal import translate_file
from tla_rust_b200.front.report import format_result
from oracle.tlc_oracle import Oracle
from test_frontend import README_BUGGY

README_TRACE = [
    dict(bob_account=10, money=(1, 10), alice_account=10, pc=("Transfer", "Transfer"), account_total=20),
    dict(bob_account=10, money=(1, 10), alice_account=10, pc=("A", "Transfer"), account_total=20),
    dict(bob_account=10, money=(1, 10), alice_account=10, pc=("A", "A"), account_total=20),
    dict(bob_account=10, money=(1, 10), alice_account=9, pc=("B", "A"), account_total=20),
    dict(bob_account=11, money=(1, 10), alice_account=9, pc=("C", "A"), account_total=20),
    dict(bob_account=11, money=(1, 10), alice_account=-1, pc=("C", "B"), account_total=20),
]
README_ACTIONS = [None, (35, 19, 40, 42), (35, 19, 40, 42), (42, 12, 45, 63), (47, 12, 50, 65), (42, 12, 45, 63)]


def _copy(name, edits=(), cfg=None):
    d = tempfile.mkdtemp(prefix="tlag_t_")
    src = open(os.path.join(REF, name + ".tla")).read()
    for a, b in edits:
        src = src.replace(a, b)
    p = os.path.join(d, name + ".tla")
    open(p, "w").write(src)
    if cfg is not None:
        open(os.path.join(d, name + ".cfg"), "w").write(cfg)
    elif os.path.exists(os.path.join(REF, name + ".cfg")):
        open(os.path.join(d, name + ".cfg"), "w").write(open(os.path.join(REF, name + ".cfg")).read())
    translate_file(p)
    return p


@needs_reference
def test_readme_transcript_exact():
    p = _copy("pcal_intro", README_BUGGY, cfg="SPECIFICATION Spec\n")
    m = Model(p)
    r = Oracle(m).run()
    assert r.verdict == "assert"
    assert r.error_text == "Failure of assertion at line 16, column 4."
    assert (r.generated, r.distinct, r.queue, r.depth) == (9097, 6164, 999, 7)      # README.md:319-320
    assert [st for st, _ in r.trace] == README_TRACE                               # README.md:271-311
    assert [None if a is None else a[2] for _, a in r.trace] == README_ACTIONS      # README.md:278-306
    txt = format_result(r, m.vars, m.module_name)
    assert "State 6: <Action line 42, col 12 to line 45, col 63 of module pcal_intro>" in txt
    assert txt.rstrip().endswith("The depth of the complete state graph search is 7.")
    assert "9097 states generated, 6164 distinct states found, 999 states left on queue." in txt


@needs_reference
def test_bundled_specs_no_error():
    p = _copy("pcal_intro")                       # README.md:349-352 "should produce no errors"
    r = Oracle(Model(p)).run()
    assert (r.verdict, r.generated, r.distinct, r.depth, r.init_states) == ("ok", 5850, 3800, 5, 400)
    p = _copy("atomic_add")
    r = Oracle(Model(p)).run()
    assert (r.verdict, r.generated, r.distinct, r.depth) == ("ok", 7, 5, 4)
    assert os.path.exists(os.path.splitext(p)[0] + ".cfg") and os.path.exists(os.path.splitext(p)[0] + ".old")


@needs_reference
def test_small_corpus_counts():
    ex = REF + "/examples/SpecifyingSystems/"
    r = Oracle(Model(ex + "HourClock/HourClock.tla")).run()
    assert (r.verdict, r.distinct) == ("ok", 12)            # HourClock.tla:4-5
    r = Oracle(Model(REF + "/examples/Paxos/MCPaxos.tla")).run()
    assert (r.verdict, r.generated, r.distinct, r.depth) == ("ok", 82, 25, 9)
    m = Model(REF + "/examples/Paxos/MCConsensus.tla")
    m.check_deadlock = False
    r = Oracle(m).run()
    assert (r.verdict, r.distinct, r.init_states) == ("ok", 4, 4)


@needs_reference
def test_oracle_handles_raft_and_ssi_at_small_bounds():
    """BASELINE configs #4/#5 (raft.tla, serializableSnapshotIsolation.tla) through the front end and O1 at reduced
    bounds.  The reference holds no counts for them (parity unpinned at the TLC boundary); these pin
    the oracle against itself: RECURSIVE operators, LAMBDA arguments, CHOOSE, bags as functions (raft.tla:117-135),
    @@ / :>, SelectSeq, CONSTRAINT semantics."""
    from conftest import ROOT
    m = Model(ROOT + "/models/MCssi.tla", extra_dirs=[REF + "/examples"])
    r = Oracle(m).run()
    assert (r.verdict, r.generated, r.distinct, r.depth) == ("ok", 945, 569, 9)
    m = Model(ROOT + "/models/MCraft.tla", extra_dirs=[REF + "/examples"])
    m.check_deadlock = False
    r = Oracle(m).run()
    assert (r.verdict, r.generated, r.distinct, r.depth) == ("ok", 6185, 694, 12)


# ---- the reference's SECOND known-answer transcript: AdvancedExamples/testout2 (TLC 1.57 on MCInnerSerial) ---------
def _testout2_numbers():
    import re
    txt = open(REF + "/examples/SpecifyingSystems/AdvancedExamples/testout2").read()
    m1 = re.search(r"Finished computing initial states: (\d+) distinct states generated", txt)
    m2 = re.search(r"^(\d+) states generated, (\d+) distinct states found, (\d+) states left on queue\.\s*$", txt, re.M)
    m3 = re.search(r"The state graph has diameter (\d+)\.", txt)
    return int(m1.group(1)), int(m2.group(1)), int(m2.group(2)), int(m2.group(3)), int(m3.group(1))


@needs_reference
def test_testout2_initial_states_match_the_front_end():
    """`Finished computing initial states: 4 distinct states generated.` (testout2:3)"""
    from tla_rust_b200.front.spec import Model
    m = Model(ROOT + "/models/MCInnerSerialTyped.tla",
              extra_dirs=[REF + "/examples/SpecifyingSystems/AdvancedExamples"])
    assert len(m.initial_states()) == _testout2_numbers()[0] == 4


@needs_reference
def test_testout2_final_counts_match_the_recorded_bytecode_run():
    """TLC: `6181 states generated, 195 distinct states found, 0 states left on queue.  The state graph has diameter
    5.` (22 h of CPU in 2001).  The AST oracle O1 cannot finish this model (a successor costs it minutes); the compiled
    model on the CPU bytecode engine O2 takes ~7 min on 8 cores, so its result is recorded in the fixture
    (tests/golden/make_golden.py) and re-run only with TLAG_SLOW=1."""
    import os
    from tla_rust_b200.compiled import load_compiled
    _init, gen, dist, queue, diam = _testout2_numbers()
    cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", "MCInnerSerial.tlagz"))
    o2 = exp["o2"]
    assert (o2["verdict"], o2["generated"], o2["distinct"], o2["depth"]) == (0, gen, dist, diam) == (0, 6181, 195, 5)
    assert queue == 0 and sum(o2["levels"]) == dist and o2["levels"][0] == 4
    if os.environ.get("TLAG_SLOW"):
        from oracle import cpu_engine
        r = cpu_engine.run(cm, init, n_threads=os.cpu_count() or 1, deadlock=info["deadlock"])
        assert (r["generated"], r["distinct"], r["depth"], r["fp_xor"]) == (gen, dist, diam, o2["fp_xor"])


def test_readme_transcript_counts_from_the_bytecode_engine():
    """README.md:319-320 again, this time from the COMPILED model: the CPU bytecode engine in its sequential mode
    (one worker, FIFO order, stop at the first Assert failure) prints TLC's error-time numbers -- 9097 states
    generated, 6164 distinct, 999 left on queue, depth 7.  This pins the compiler's successor order and the
    initial-state order, not only the set of reachable states."""
    import os
    from tla_rust_b200.compiled import load_compiled
    from oracle import cpu_engine
    cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", "pcal_intro_readme_buggy.tlagz"))
    r = cpu_engine.run(cm, init, exact=True)
    assert (r["verdict"], r["generated"], r["distinct"], r["queue"], r["depth"]) == (2, 9097, 6164, 999, 7)
    assert cm.asserts[r["detail"]][0] == "Failure of assertion at line 16, column 4."
