"""Time the CUDA engine on committed fixtures (tests/golden/*.tlagz): whole-BFS device time and distinct
states/s, with the counts checked against the recorded oracle result.  Used for the raft numbers in
DESIGN.md (the contract bench, bench.py, stays on MCPaxos3_b4).

    python tools/fixture_bench.py MCraft_s3_m [MCraft_s3_l ...] [--reps 3]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tla_rust_b200.compiled import load_compiled  # noqa: E402
from tla_rust_b200.engine import Engine  # noqa: E402


def main():
    args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and sys.argv[i - 1] != "--reps"]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    for name in args:
        cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", name + ".tlagz"))
        o2 = exp["o2"]
        e = Engine(cm, deadlock=info["deadlock"], native="--native" in sys.argv)
        best = None
        for r in range(reps):
            t0 = time.time()
            if r:
                e.restart()          # clears the table and re-seeds the retained initial states
            else:
                e.seed(init)
            if "--waves" in sys.argv:
                while True:
                    ws = e.step()
                    if ws["expanded"]:
                        print(json.dumps({"fixture": name, "rep": r, "level": ws["level"], "expanded": ws["expanded"],
                                          "generated": ws["generated"], "kernel_ms": round(ws["kernel_ms"], 3)}), flush=True)
                    if ws["verdict"] != 5:
                        break
                res = e.result()
            else:
                res = e.run()
            wall = time.time() - t0
            ok = (res["generated"], res["distinct"], res["depth"]) == (o2["generated"], o2["distinct"], o2["depth"])
            dev = res["device_seconds"]
            if best is None or dev < best[0]:
                best = (dev, wall)
            if not ok:
                print(json.dumps({"fixture": name, "error": "count mismatch", "got": res, "want": o2}))
                break
        print(json.dumps({"fixture": name, "W": cm.W, "code_len": int(len(cm.code)), "frame_words": cm.frame_words,
                          "distinct": res["distinct"], "generated": res["generated"], "depth": res["depth"],
                          "device_s": round(best[0], 4), "wall_s": round(best[1], 4),
                          "distinct_per_s": round(res["distinct"] / best[0]), "generated_per_s": round(res["generated"] / best[0]),
                          "counts_match_oracle": ok, "launches": e.launches(),
                          "native": "--native" in sys.argv,
                          "digest_matches_oracle": tuple(e.digest()) == (o2["fp_xor"], o2["fp_sum"])}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
