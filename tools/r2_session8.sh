#!/bin/bash
# Round 2, GPU session 8 (2 GPUs): validation of the ownership probe (k = W for SSI), parked exchange buffers (end-to-end
# leg at N = 2), host-driven path accounting; 2-GPU parity suite once more.  Log tail to stdout.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s8.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step "pytest 2-GPU parity" timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short
step "dist_check ssi 2x2 p2p (owner probe)" timeout 300 $TR --master-port 29531 tools/dist_check.py MCssi_2x2 p2p sliced
step "bench N=2 b4" timeout 400 $TR --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 2 --no-k1 --no-cpu
step "bench N=2 ssi 4x3" timeout 400 $TR --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 1 --no-k1 --no-cpu --workload MCssi_4x3
step "pytest keep_going" timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k keep_going --tb=short
grep -E "^===|^rc=|passed|failed|Error|assert" "$LOG" | cut -c1-300
grep '"metric"' "$LOG" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print(d['config']['workload'][:14], 'N', d['n_gpus'], 'value %.1f M'%(d['value']/1e6), 'ms/step', d['ms_per_step'], 'kernel_s', d['roofline']['kernel_s_per_step'], 'e2e %.1f M'%(d['e2e']['value']/1e6), d.get('exchange','')[:12])
"
grep -E "owner_words" "$LOG" | cut -c1-400
