"""Fixtures for BASELINE configs #4 and #5 at their stated bounds (run in the build container: needs /root/reference).

  MCssi_4x3   serializableSnapshotIsolation.tla, 4 transactions x 3 keys, all eight invariants, deadlock ON.  The state
              space does not end in any practical sense (levels grow ~9x: 1, 4, 32, 264, 2532, 24576, 236844, 2189052,
              ...), so the fixture records ORACLE O2's depth-bounded prefix: cumulative counts and fingerprint digests
              after every level up to level 10 (>= 10^8 distinct states, BASELINE config #5's size).
  MCraft_t4l3 raft.tla, 3 servers, MaxTerm 4, MaxLogLen 3 (BASELINE config #4) with the message bag bounded to 3
              distinct messages -- raft.tla:471 (DuplicateMessage) makes the space infinite without such a bound.

    python tests/golden/make_big.py [ssi|raft] [max_levels]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tla_rust_b200.front.spec import Model  # noqa: E402
from tla_rust_b200.checker import compile_model, encode_states  # noqa: E402
from tla_rust_b200.compiled import save_compiled  # noqa: E402
from oracle import cpu_engine  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def ssi(max_levels):
    cfg = open(ROOT + "/models/MCssi.cfg").read().replace("TxnId = {T1, T2}", "TxnId = {T1, T2, T3, T4}") \
        .replace("Key = {K1}", "Key = {K1, K2, K3}")
    m = Model(ROOT + "/models/MCssi.tla", extra_dirs=[REF + "/examples"], cfg_text=cfg)
    m.check_deadlock = True
    return "MCssi_4x3", m, dict(seq_cap=24, subroutines=True), max_levels, 1 << 28


def raft(max_levels):
    cfg = open(ROOT + "/models/MCraft_s3_l.cfg").read().replace("MaxTerm = 3", "MaxTerm = 4") \
        .replace("MaxLogLen = 2", "MaxLogLen = 3").replace("MaxMessages = 4", "MaxMessages = 3")
    m = Model(ROOT + "/models/MCraft.tla", extra_dirs=[REF + "/examples"], cfg_text=cfg)
    m.check_deadlock = True
    return "MCraft_t4l3", m, {}, max_levels, 1 << 24


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "ssi"
    max_levels = int(sys.argv[2]) if len(sys.argv) > 2 else (10 if which == "ssi" else 0)
    name, m, ckw, max_levels, cap = (ssi if which == "ssi" else raft)(max_levels)
    t0 = time.time()
    m.check_assumes()
    init = m.initial_states()
    cm = compile_model(m, init, **ckw)
    iw = encode_states(cm, init)
    print(f"{name}: compiled in {time.time() - t0:.1f}s W={cm.W} code={len(cm.code)} frame={cm.frame_words}", flush=True)
    o2 = cpu_engine.run(cm, iw, n_threads=os.cpu_count() or 1, deadlock=True, max_states=cap, max_levels=max_levels)
    exp = {"o2": {k: o2[k] for k in ("verdict", "detail", "generated", "distinct", "depth", "init_states", "fp_xor",
                                     "fp_sum", "levels", "state_idx", "level_digests")}}
    exp["o2"]["max_levels"] = max_levels
    info = {"source": f"<repo>/models/{'MCssi' if which == 'ssi' else 'MCraft'}.tla", "deadlock": True,
            "code_len": int(len(cm.code)), "W": cm.W, "o2_seconds": o2["seconds"], "o2_threads": os.cpu_count()}
    save_compiled(os.path.join(OUT, name + ".tlagz"), cm, iw, exp, info)
    print(name, {k: o2[k] for k in ("verdict", "generated", "distinct", "depth", "seconds")}, o2["levels"], flush=True)


if __name__ == "__main__":
    main()
