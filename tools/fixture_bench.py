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
    args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and sys.argv[i - 1] not in ("--reps", "--max-levels")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    max_levels = int(sys.argv[sys.argv.index("--max-levels") + 1]) if "--max-levels" in sys.argv else 0
    native = "sliced" if "--sliced" in sys.argv else ("--native" in sys.argv)
    for name in args:
        cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", name + ".tlagz"))
        o2 = exp["o2"]
        ml = max_levels or o2.get("max_levels", 0)
        if ml:     # depth-bounded prefix: expected values are the oracle's cumulative record after level `ml`
            x, sm, gen = o2["level_digests"][ml - 1]
            o2 = dict(o2, generated=gen, distinct=sum(o2["levels"][:ml]), depth=ml, fp_xor=x, fp_sum=sm)
        e = Engine(cm, deadlock=info["deadlock"], native=native)
        best = None
        for r in range(reps):
            t0 = time.time()
            if r:
                e.restart()          # clears the table and re-seeds the retained initial states
            else:
                e.seed(init)
            if "--waves" in sys.argv or ml:
                lv = 1
                while not ml or lv < ml:
                    ws = e.step()
                    lv += 1
                    if "--waves" not in sys.argv:
                        if ws["verdict"] != 5:
                            break
                        continue
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
                          "native": native, "max_levels": ml,
                          "digest_matches_oracle": tuple(e.digest()) == (o2["fp_xor"], o2["fp_sum"])}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
