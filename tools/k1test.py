import sys, time
sys.path.insert(0, '.')
import torch, numpy as np
from tla_rust_b200.engine import Engine, ProbeOnlyModel
W = 20
for logn, tl in ((20, 22), (24, 26), (26, 28), (27, 28)):
    n = 1 << logn
    g = torch.Generator(device='cuda').manual_seed(1)
    half = torch.randint(-2**31, 2**31 - 1, (n // 2, W), dtype=torch.int32, device='cuda', generator=g)
    perm = torch.randint(0, n // 2, (n // 2,), device='cuda', generator=g)
    states = torch.cat([half, half[perm]]); del half, perm
    flags = torch.zeros(n, dtype=torch.uint8, device='cuda')
    e = Engine(ProbeOnlyModel(W), table_log2=tl, native=False)
    for it in range(3):
        e.reset_table()
        t = time.time()
        ms = e.probe_batch_device(states.data_ptr(), n, flags.data_ptr())
        print(logn, tl, it, 'kernel ms', round(ms, 3), 'wall', round(time.time() - t, 3), 'new', int(flags.sum().item()), flush=True)
    e.close(); del states, flags; torch.cuda.empty_cache()
