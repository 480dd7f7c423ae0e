"""One multi-GPU BFS of a committed fixture under torchrun; rank 0 prints a JSON line with the global counts, the
XOR / SUM digest of all shards and (on a violation) the behaviour stitched across ranks.  Used by tests/test_dist_gpu.py.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
        tools/dist_check.py MCPaxos3_b2 p2p sliced [cap_records] [chunk_states] [max_levels]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tla_rust_b200.compiled import load_compiled  # noqa: E402
from tla_rust_b200.dist import DistributedBFS  # noqa: E402
from tla_rust_b200.engine import Engine  # noqa: E402


def main():
    name, exchange, engine = sys.argv[1:4]
    cap = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 22
    chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 20
    max_levels = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cm, init, exp, info = load_compiled(os.path.join(ROOT, "tests", "golden", name + ".tlagz"))
    e = Engine(cm, deadlock=info["deadlock"], device=local, native="sliced" if engine == "sliced" else False)
    d = DistributedBFS(e, cm, rank, world, f"cuda:{local}", cap_records=cap, chunk_states=chunk, exchange=exchange)
    d.seed(init)
    out = d.run(max_levels=(max_levels - 1) if max_levels else 1 << 20)
    dig = d.global_digest()
    cex = d.counterexample()
    rt = torch.tensor([getattr(d, "retries", 0)], dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(rt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"verdict": out["verdict"], "generated": out["generated"], "distinct": out["distinct"],
                          "depth": out["depth"], "local_distinct": out["local"]["distinct"], "digest": list(dig),
                          "exchange": d.exchange, "exchange_note": d.exchange_note, "exchanges": d.exchanges,
                          "cap_records": d.cap_records, "owner_words": d.owner_words, "owner_skew_k2": getattr(d, "owner_skew", None), "retries": getattr(d, "retries", 0), "retries_max": int(rt.item()), "chunk_states": d.chunk_states,
                          "cex": None if cex is None else {"verdict": cex[0], "detail": cex[1], "states": cex[2].tolist(),
                                                           "actions": cex[3].tolist()}}))
    e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
