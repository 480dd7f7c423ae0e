"""Regenerates the committed parity fixtures under tests/golden/ (run in the build container,
where /root/reference exists; the GPU box only sees the generated .tlagz files).

For every model: translate PlusCal if needed, compile to bytecode, run ORACLE O1 (Python AST
evaluator, pinned by README.md:267-321) and ORACLE O2 (C bytecode engine) and store their
counts, per-level sizes and fingerprint digests as the expected values.

    python tests/golden/make_golden.py [--only name]
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tla_rust_b200.front.spec import Model  # noqa: E402
from tla_rust_b200.front.pcal import translate_file  # noqa: E402
from tla_rust_b200.checker import compile_model, encode_states  # noqa: E402
from tla_rust_b200.compiled import save_compiled  # noqa: E402
from oracle.tlc_oracle import Oracle  # noqa: E402
from oracle import cpu_engine  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
README_BUGGY = (("     alice_account := alice_account - money;", "     A: alice_account := alice_account - money;"),
                ("     bob_account := bob_account + money;", "     B: bob_account := bob_account + money;"))


def pcal_copy(name, edits=(), cfg=None):
    d = tempfile.mkdtemp(prefix="tlag_gold_")
    src = open(os.path.join(REF, name + ".tla")).read()
    for a, b in edits:
        assert a in src
        src = src.replace(a, b)
    p = os.path.join(d, name + ".tla")
    open(p, "w").write(src)
    if cfg is not None:
        open(os.path.join(d, name + ".cfg"), "w").write(cfg)
    elif os.path.exists(os.path.join(REF, name + ".cfg")):
        shutil.copy(os.path.join(REF, name + ".cfg"), d)
    translate_file(p)
    return p


def demo_copy(name):
    d = tempfile.mkdtemp(prefix="tlag_gold_")
    for ext in (".tla", ".cfg"):
        shutil.copy(os.path.join(ROOT, "models", "demo", name + ext), d)
    p = os.path.join(d, name + ".tla")
    translate_file(p)
    return p


MODELS = {
    # name: (builder -> (tla path, Model kwargs), deadlock check, run O1?)
    "atomic_add": (lambda: (pcal_copy("atomic_add"), {}), True, True),
    "pcal_intro": (lambda: (pcal_copy("pcal_intro"), {}), True, True),
    "pcal_intro_readme_buggy": (lambda: (pcal_copy("pcal_intro", README_BUGGY, cfg="SPECIFICATION Spec\n"), {}), True, True),
    "MCPaxos": (lambda: (REF + "/examples/Paxos/MCPaxos.tla", {}), True, True),
    "MCPaxos3": (lambda: (ROOT + "/models/MCPaxos3.tla", {"extra_dirs": [REF + "/examples/Paxos"]}), True, True),
    "MCPaxos3_sym": (lambda: (ROOT + "/models/MCPaxos3.tla",
                              {"extra_dirs": [REF + "/examples/Paxos"], "cfg_path": ROOT + "/models/MCPaxos3_sym.cfg"}),
                     True, True),
    "MCPaxos3_b2": (lambda: (ROOT + "/models/MCPaxos3.tla",
                             {"extra_dirs": [REF + "/examples/Paxos"],
                              "cfg_text": open(ROOT + "/models/MCPaxos3.cfg").read().replace("MaxBallot = 1", "MaxBallot = 2")}),
                    True, False),
    "MCVoting": (lambda: (REF + "/examples/Paxos/MCVoting.tla", {}), False, True),
    # verdict kinds other than "ok": deadlock (Voting terminates: MCVoting with deadlock checking ON) and an
    # invariant violation with a counterexample (builder-authored lost-update demo, models/demo/race.tla)
    "MCVoting_deadlock": (lambda: (REF + "/examples/Paxos/MCVoting.tla", {}), True, True),
    "demo_race": (lambda: (demo_copy("race"), {}), True, True),
    "demo_lock": (lambda: (demo_copy("lock"), {}), True, True),
    # bounded sequences (Seq(S) with a capacity = constraint bound + 1: states one element past the CONSTRAINT
    # are generated and counted, then not explored -- FIFO/MCInnerFIFO.cfg:23-31)
    "MCInnerFIFO": (lambda: (REF + "/examples/SpecifyingSystems/FIFO/MCInnerFIFO.tla", {}), True, True, 4),
    "MCAlternatingBit": (lambda: (REF + "/examples/SpecifyingSystems/TLC/MCAlternatingBit.tla", {}), True, True, 4),
    # BASELINE config #4 (examples/raft.tla): sparse containers for the message bag and the history variables
    "MCraft": (lambda: (ROOT + "/models/MCraft.tla", {"extra_dirs": [REF + "/examples"]}), True, True),
    "MCraft_s3": (lambda: (ROOT + "/models/MCraft.tla",
                           {"extra_dirs": [REF + "/examples"], "cfg_path": ROOT + "/models/MCraft_s3.cfg"}), True, True),
    "MCraft_s3_m": (lambda: (ROOT + "/models/MCraft.tla",
                             {"extra_dirs": [REF + "/examples"], "cfg_path": ROOT + "/models/MCraft_s3_m.cfg"}), True, False),
    "MCraft_s3_l": (lambda: (ROOT + "/models/MCraft.tla",
                             {"extra_dirs": [REF + "/examples"], "cfg_path": ROOT + "/models/MCraft_s3_l.cfg"}), True, False),
    # BASELINE config #5 at its smallest bounds (2 transactions x 1 key), all eight invariants: operator subroutines
    "MCssi": (lambda: (ROOT + "/models/MCssi.tla", {"extra_dirs": [REF + "/examples"]}), True, True, 8, {"subroutines": True}),
    # ... 3 transactions x 1 key (frame 2910 words; O1 ran once: 152554 / 90430 / 13, tests/test_containers.py), 2 x 2
    # (sequence capacity 16, frame 3400 words; O1: 50121 / 29629 / 13) and the same state space with capacity 24
    # (wider packed states, frame 5080 words: the engine's 8192-word frame class)
    "MCssi_3x1": (lambda: (ROOT + "/models/MCssi.tla",
                           {"extra_dirs": [REF + "/examples"],
                            "cfg_text": open(ROOT + "/models/MCssi.cfg").read().replace("TxnId = {T1, T2}", "TxnId = {T1, T2, T3}")}),
                  True, False, 13, {"subroutines": True}),
    "MCssi_2x2": (lambda: (ROOT + "/models/MCssi.tla",
                           {"extra_dirs": [REF + "/examples"],
                            "cfg_text": open(ROOT + "/models/MCssi.cfg").read().replace("Key = {K1}", "Key = {K1, K2}")}),
                  True, False, 16, {"subroutines": True}),
    "MCssi_2x2_wide": (lambda: (ROOT + "/models/MCssi.tla",
                                {"extra_dirs": [REF + "/examples"],
                                 "cfg_text": open(ROOT + "/models/MCssi.cfg").read().replace("Key = {K1}", "Key = {K1, K2}")}),
                       True, False, 24, {"subroutines": True}),
    # AdvancedExamples/MCInnerSerial: the reference's second TLC transcript (testout2: 6181 generated / 195 distinct /
    # diameter 5, 22 h of CPU in 2001).  O1 cannot finish it; O2 takes ~7 min on 8 cores and reproduces it exactly.
    "MCInnerSerial": (lambda: (ROOT + "/models/MCInnerSerialTyped.tla",
                               {"extra_dirs": [REF + "/examples/SpecifyingSystems/AdvancedExamples"]}),
                      True, False, None, {"type_hint": "TypeOK"}),
    "Containers": (lambda: (ROOT + "/tests/specs/Containers.tla", {}), False, True),
    "HourClock": (lambda: (REF + "/examples/SpecifyingSystems/HourClock/HourClock.tla", {}), True, True),
    "AsynchInterface": (lambda: (REF + "/examples/SpecifyingSystems/AsynchronousInterface/AsynchInterface.tla", {}), True, True),
}


def main():
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    for name, spec in MODELS.items():
        mk, deadlock, run_o1 = spec[:3]
        seq_cap = spec[3] if len(spec) > 3 else None
        ckw = spec[4] if len(spec) > 4 else {}
        if only and name != only:
            continue
        t0 = time.time()
        path, kw = mk()
        m = Model(path, **kw)
        m.check_deadlock = deadlock
        m.check_assumes()
        init = m.initial_states()
        cm = compile_model(m, init, seq_cap=seq_cap, **ckw)
        iw = encode_states(cm, init)
        o2 = cpu_engine.run(cm, iw, n_threads=os.cpu_count() or 1, deadlock=deadlock, max_states=1 << 27)
        exp = {"o2": {k: o2[k] for k in ("verdict", "detail", "generated", "distinct", "depth", "init_states",
                                         "fp_xor", "fp_sum", "levels", "state_idx")}}
        if run_o1:
            r = Oracle(m).run()
            exp["o1"] = dict(r.summary())
            if r.verdict == "ok":
                assert (r.generated, r.distinct, r.depth) == (o2["generated"], o2["distinct"], o2["depth"]), \
                    (name, r.summary(), exp["o2"])
        info = {"source": path.replace(REF, "<reference>").replace(ROOT, "<repo>"), "deadlock": deadlock,
                "code_len": int(len(cm.code)), "W": cm.W}
        save_compiled(os.path.join(OUT, name + ".tlagz"), cm, iw, exp, info)
        print(f"{name}: W={cm.W} code={len(cm.code)} o2={ {k: o2[k] for k in ('verdict','generated','distinct','depth')} } "
              f"o1={exp.get('o1')} ({time.time()-t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
