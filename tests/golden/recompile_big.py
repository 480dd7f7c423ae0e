"""Re-lowers the large Paxos fixtures (MCPaxos3_b3 / _b4) after a compiler or ISA change while keeping the
expected counts recorded by the full ORACLE O2 runs (20 s and 11 min of CPU respectively; see
make_golden.py for how the small fixtures and the original runs are produced)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tla_rust_b200.front.spec import Model  # noqa: E402
from tla_rust_b200.checker import compile_model, encode_states  # noqa: E402
from tla_rust_b200.compiled import save_compiled, load_compiled  # noqa: E402

REF = "/root/reference/examples/Paxos"
for mb in (3, 4):
    path = os.path.join(ROOT, "tests", "golden", f"MCPaxos3_b{mb}.tlagz")
    _, _, exp, info = load_compiled(path)
    cfg = open(ROOT + "/models/MCPaxos3.cfg").read().replace("MaxBallot = 1", f"MaxBallot = {mb}")
    m = Model(ROOT + "/models/MCPaxos3.tla", extra_dirs=[REF], cfg_text=cfg)
    init = m.initial_states()
    cm = compile_model(m, init)
    info.update(code_len=int(len(cm.code)), W=cm.W)
    save_compiled(path, cm, encode_states(cm, init), exp, info)
    print(path, "W", cm.W, "code", len(cm.code), exp["o2"]["distinct"])
