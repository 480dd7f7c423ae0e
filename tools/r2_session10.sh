#!/bin/bash
# Round 2, GPU session 10 (8 GPUs, short): SSI 4x3 (config #5's 8-GPU stress) with the ownership probe, b4 at N = 8 with
# parked exchange buffers (end-to-end leg), raft.
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s10.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step "bench N=8 ssi 4x3" timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 1 --no-k1 --no-cpu --workload MCssi_4x3
step "bench N=8 b4" timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-k1 --no-cpu
step "bench N=8 raft t4l3" timeout 200 $TR --master-port 29523 bench.py --gpus 8 --steps 2 --warmup 1 --no-k1 --no-cpu --workload MCraft_t4l3
grep -E "^===|^rc=|Error|error" "$LOG" | cut -c1-300
grep '"metric"' "$LOG" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print(d['config']['workload'][:14], 'N', d['n_gpus'], 'value %.1f M'%(d['value']/1e6), 'ms/step', d['ms_per_step'], 'kernel_s', d['roofline']['kernel_s_per_step'], 'e2e %.1f M'%(d['e2e']['value']/1e6), d.get('exchange','')[:12], 'launches', d['gpu_launches'])
"
