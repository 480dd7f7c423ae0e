#!/bin/bash
# Round 2, GPU session 6 (8 GPUs, expensive: keep it short): the contract bench at N=8 on the peer-memory exchange, then
# BASELINE configs #5 (SSI 4x3, 168 M states) and #4 (raft MaxTerm 4 / MaxLogLen 3) on 8 GPUs, counts + digests checked.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_session6.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s6.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step "bench N=8 b4" timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 --no-k1 --no-cpu
step "bench N=8 ssi 4x3" timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 1 --warmup 1 --no-k1 --no-cpu --workload MCssi_4x3
step "bench N=8 raft t4l3" timeout 200 $TR --master-port 29523 bench.py --gpus 8 --steps 1 --warmup 1 --no-k1 --no-cpu --workload MCraft_t4l3
step "bench N=4 b4" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 2 --no-k1 --no-cpu
tail -3 "$LOG"
