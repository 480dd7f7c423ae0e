"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against the
recorded oracle results of the committed fixtures and against ORACLE O2 run live on the host."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from tla_rust_b200.compiled import load_compiled
from tla_rust_b200.checker import decode_state, result_from_engine

pytestmark = pytest.mark.gpu


def _engine(cm, **kw):
    from tla_rust_b200.engine import Engine
    kw.setdefault("native", False)          # this file pins the interpreter kernel unless a test says otherwise
    return Engine(cm, **kw)


def _check_fixture(name, native=False):
    from oracle import cpu_engine
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    e = _engine(cm, deadlock=info["deadlock"], native=native)
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
    if r["verdict"] == 2:   # which Assert site is hit first depends on the (parallel) discovery order
        assert cm.asserts[r["detail"]][0] == cm.asserts[o2["detail"]][0]
    # bit-exact state set: fingerprint digest of everything the GPU stored (checksum of checksums);
    # MCPaxos3_b4 is the BASELINE-size case (352,133,865 states): digest computed on the device
    assert e.digest() == (o2["fp_xor"], o2["fp_sum"])
    if r["distinct"] <= 1 << 20:
        states = e.read_states(0, r["distinct"])
        assert cpu_engine.digest(states, cm.W) == (o2["fp_xor"], o2["fp_sum"])
    if "o1" in exp and exp["o1"]["verdict"] == "ok":
        assert (r["generated"], r["distinct"], r["depth"]) == (exp["o1"]["generated"], exp["o1"]["distinct"],
                                                               exp["o1"]["depth"])
    assert e.launches() >= len(levels)
    e.close()


@pytest.mark.parametrize("name", ["atomic_add", "pcal_intro", "pcal_intro_readme_buggy", "MCPaxos", "MCVoting",
                                  "MCVoting_deadlock", "demo_race", "demo_lock", "MCInnerFIFO", "MCAlternatingBit",
                                  "MCPaxos3_sym", "Containers", "HourClock", "AsynchInterface", "MCPaxos3", "MCPaxos3_b2",
                                  "MCPaxos3_b3", "MCPaxos3_b4"])
def test_bfs_matches_oracle(name):
    _check_fixture(name)


@pytest.mark.parametrize("name,counts", [("MCssi", [0, 945, 569, 9]), ("MCssi_3x1", [0, 152554, 90430, 13]),
                                         ("MCssi_2x2", [0, 50121, 29629, 13]), ("MCssi_2x2_wide", [0, 50121, 29629, 13])])
def test_ssi_subroutine_model_on_device(name, counts):
    """serializableSnapshotIsolation.tla, eight invariants: 2 transactions x 1 key (frame 1720 words: 2048 class), 3 x 1
    and 2 x 2 (2910 / 3400 words: 4096 class), and 2 x 2 with a larger sequence capacity (5080 words: 8192 class); the
    counts are the ones the AST oracle O1 produced (tests/test_containers.py), the digests the CPU bytecode engine's."""
    _check_fixture(name)
    _, _, exp, _ = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    o2 = exp["o2"]
    assert [o2["verdict"], o2["generated"], o2["distinct"], o2["depth"]] == counts


def _check_prefix(name, levels, native):
    """depth-bounded run: counts and fingerprint digest after `levels` levels against the oracle's per-level record"""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, name + ".tlagz"))
    o2 = exp["o2"]
    x, sm, gen = o2["level_digests"][levels - 1]
    e = _engine(cm, deadlock=info["deadlock"], native=native)
    e.seed(init)
    for _ in range(levels - 1):
        ws = e.step()
        assert ws["verdict"] == 5
    r = e.result()
    assert (r["generated"], r["distinct"], r["depth"]) == (gen, sum(o2["levels"][:levels]), levels)
    assert e.digest() == (x, sm)
    e.close()


def test_config5_ssi_4x3_prefix_matches_oracle_on_both_engines():
    """BASELINE config #5 at its stated bound (4 transactions x 3 keys, eight invariants, deadlock ON): the first 8
    levels (2,453,305 states) on the interpreter kernel and on the sliced kernels, bit-exact against ORACLE O2's
    per-level digests (the 10-level, 168 M-state job is bench.py --workload MCssi_4x3)."""
    _check_prefix("MCssi_4x3", 8, False)
    _check_prefix("MCssi_4x3", 8, "sliced")


def test_config4_raft_t4l3_matches_oracle_on_both_engines():
    """BASELINE config #4 at its stated bound (3 servers, MaxTerm 4, MaxLogLen 3): 11,296,712 states, whole space."""
    _check_fixture("MCraft_t4l3")
    _check_fixture("MCraft_t4l3", native="sliced")


def test_assert_trace_is_a_shortest_counterexample():
    """README.md:267-316: the failing assertion is reached after 5 steps from an initial state; the GPU
    trace must be a valid 6-state behaviour ending in a state with pc = C for a process whose alice < 0."""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "pcal_intro_readme_buggy.tlagz"))
    e = _engine(cm)
    e.seed(init)
    r = e.run()
    assert r["verdict"] == 2
    states, acts = e.trace(r["state_idx"])
    res = result_from_engine(cm, r, (states, acts))
    assert res.error_text == "Failure of assertion at line 16, column 4."
    assert len(res.trace) == 6 and res.trace[0][1] is None
    last = res.trace[-1][0]
    assert last["alice_account"] < 0 and "C" in last["pc"]
    assert res.trace[0][0]["alice_account"] == 10 and res.trace[0][0]["pc"] == ("Transfer", "Transfer")
    e.close()


def test_exact_replay_reproduces_the_readme_transcript_on_the_device():
    """README.md:267-321, the reference's only known-answer transcript for this path, on the GPU: TLAG_F_EXACT runs the
    search as one sequential worker (one warp, one state at a time, FIFO order, stop AT the failed assert) -- the counts
    TLC printed (9097 generated / 6164 distinct / 999 on queue / depth 7) and the 6-state trace it printed."""
    from tla_rust_b200.front.report import format_result
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "pcal_intro_readme_buggy.tlagz"))
    e = _engine(cm, exact=True)
    e.seed(init)
    r = e.run()
    assert r["verdict"] == 2
    assert (r["generated"], r["distinct"], r["queue_left"], r["depth"]) == (9097, 6164, 999, 7)       # README.md:319-320
    res = result_from_engine(cm, r, e.trace(r["state_idx"]))
    want = [dict(bob_account=10, money=(1, 10), alice_account=10, pc=("Transfer", "Transfer"), account_total=20),
            dict(bob_account=10, money=(1, 10), alice_account=10, pc=("A", "Transfer"), account_total=20),
            dict(bob_account=10, money=(1, 10), alice_account=10, pc=("A", "A"), account_total=20),
            dict(bob_account=10, money=(1, 10), alice_account=9, pc=("B", "A"), account_total=20),
            dict(bob_account=11, money=(1, 10), alice_account=9, pc=("C", "A"), account_total=20),
            dict(bob_account=11, money=(1, 10), alice_account=-1, pc=("C", "B"), account_total=20)]
    assert [st for st, _ in res.trace] == want                                                          # README.md:271-311
    assert [None if a is None else a[2] for _, a in res.trace] == [None, (35, 19, 40, 42), (35, 19, 40, 42), (42, 12, 45, 63),
                                                                   (47, 12, 50, 65), (42, 12, 45, 63)]   # README.md:278-306
    txt = format_result(res, cm.vars, "pcal_intro")
    assert "Failure of assertion at line 16, column 4." in txt
    assert "9097 states generated, 6164 distinct states found, 999 states left on queue." in txt
    assert txt.rstrip().endswith("The depth of the complete state graph search is 7.")
    e.close()
    # the same flag on a deadlocking model: stops at the first state without successors, in FIFO order
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "MCVoting_deadlock.tlagz"))
    from oracle import cpu_engine
    o = cpu_engine.run(cm, init, deadlock=True, exact=True)
    e = _engine(cm, deadlock=True, exact=True)
    e.seed(init)
    r = e.run()
    assert (r["verdict"], r["generated"], r["distinct"], r["queue_left"], r["state_idx"]) == (
        3, o["generated"], o["distinct"], o["queue"], o["state_idx"])
    e.close()


def test_invariant_and_deadlock_traces_are_valid_counterexamples():
    """Verdict kinds other than ok: the reported state really is an error state and the parent chain is a
    behaviour of minimal length starting in an initial state (level-synchronous BFS => shortest)."""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "demo_race.tlagz"))
    e = _engine(cm)
    e.seed(init)
    r = e.run()
    assert r["verdict"] == 1 and cm.invariants[r["detail"]] == "Correct"
    res = result_from_engine(cm, r, e.trace(r["state_idx"]))
    last = res.trace[-1][0]
    assert last["pc"] == ("Done", "Done") and last["counter"] == 1      # lost update
    assert len(res.trace) == 5 and res.trace[0][1] is None and res.trace[0][0]["counter"] == 0
    for (s0, _), (s1, a1) in zip(res.trace, res.trace[1:]):
        assert a1 is not None and s0 != s1
    e.close()
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "MCVoting_deadlock.tlagz"))
    e = _engine(cm, deadlock=True)
    e.seed(init)
    r = e.run()
    assert r["verdict"] == 3 and r["depth"] == exp["o2"]["depth"]
    states, acts = e.trace(r["state_idx"])
    assert len(states) >= 2 and acts[0] == -1 and (states[0] == init[0]).all()
    e.close()


def test_keep_going_explores_the_whole_space_and_reports_the_first_violation():
    """TLAG_F_KEEP_GOING (TLC's -continue): the search does not stop at the violation; the verdict, the violating state
    and its trace are those of the FIRST violation, the counts are those of the whole reachable space."""
    cm, init, exp, info = load_compiled(os.path.join(GOLDEN, "demo_race.tlagz"))
    e = _engine(cm)
    e.seed(init)
    first = e.run()
    e.close()
    e = _engine(cm, keep_going=True)
    e.seed(init)
    r = e.run()
    assert (r["verdict"], r["detail"], r["state_idx"]) == (first["verdict"], first["detail"], first["state_idx"]) and r["verdict"] == 1
    assert r["distinct"] >= first["distinct"] and r["generated"] >= first["generated"] and r["queue_left"] == 0
    res = result_from_engine(cm, r, e.trace(r["state_idx"]))
    assert res.trace[-1][0]["counter"] == 1 and res.trace[-1][0]["pc"] == ("Done", "Done")
    e.close()


@pytest.mark.parametrize("W,n", [(1, 1000), (3, 5000), (4, 100000), (20, 200000), (7, 0), (64, 300)])
def test_probe_batch_matches_cpu(W, n):
    """K1 alone through the C ABI: exactly one 'new' flag per distinct state, same set as the CPU oracle;
    edge cases: empty batch, widest state, ragged (non-vectorisable) widths."""
    from tla_rust_b200.engine import Engine, ProbeOnlyModel
    from oracle import cpu_engine
    rng = np.random.default_rng(1234 + W)
    base = rng.integers(0, 2**32, size=(max(n // 2, 1), W), dtype=np.uint64).astype(np.uint32)
    states = np.concatenate([base, base[rng.integers(0, len(base), size=n - len(base))]]) if n else base[:0]
    e = Engine(ProbeOnlyModel(W), table_log2=20)
    flags = e.probe_batch(states)
    assert flags.shape == (n,)
    if n:
        cflags, _ = cpu_engine.probe_batch(states, W, 20, 1)
        uniq, inv = np.unique(states, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        assert int(flags.sum()) == len(uniq) == int(cflags.sum())
        per = np.bincount(inv, weights=flags, minlength=len(uniq))
        assert (per == 1).all()
        # idempotence: probing again finds nothing new
        assert int(e.probe_batch(states).sum()) == 0
    e.close()


def test_probe_roundtrip_full_size_properties():
    """BASELINE-size property test (SURVEY §8d synthetic K1 input, scaled to 2^24 here): #new == #distinct,
    second pass == 0, after reset the same batch is all-new again."""
    import torch
    from tla_rust_b200.engine import Engine, ProbeOnlyModel
    W, n = 20, 1 << 24
    g = torch.Generator(device="cuda").manual_seed(0x5EED)
    half = torch.randint(0, 2**31 - 1, (n // 2, W), dtype=torch.int32, device="cuda", generator=g)
    perm = torch.randint(0, n // 2, (n // 2,), device="cuda", generator=g)
    states = torch.cat([half, half[perm]])
    flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
    e = Engine(ProbeOnlyModel(W), table_log2=26)
    e.probe_batch_device(states.data_ptr(), n, flags.data_ptr())
    torch.cuda.synchronize()
    n_new = int(flags.sum().item())
    assert n_new == n // 2          # random 80-byte rows: distinct with overwhelming probability
    e.probe_batch_device(states.data_ptr(), n, flags.data_ptr())
    assert int(flags.sum().item()) == 0
    e.reset_table()
    e.probe_batch_device(states.data_ptr(), n, flags.data_ptr())
    assert int(flags.sum().item()) == n // 2
    e.close()


@pytest.mark.parametrize("name", ["MCraft", "MCraft_s3", "MCraft_s3_m", "MCraft_s3_l"])
def test_raft_fixtures_match_oracle(name):
    _check_fixture(name)
