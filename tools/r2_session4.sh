#!/bin/bash
# Round 2, GPU session 4 (2 GPUs): after the retry / rollback / local-owner fixes -- 2-GPU parity suite, bench N=2.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_session4.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
LOG="$OUT/r2_s4.log"
: > "$LOG"
export TLAG_NO_BUILD=1
step() { echo "=== $1" | tee -a "$LOG"; shift; ( "$@" ) >> "$LOG" 2>&1; echo "rc=$?" | tee -a "$LOG"; }
step "pytest 2-GPU parity" timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q
step "bench N=2 p2p" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-k1
step "bench N=1 (new defaults)" timeout 900 python bench.py --steps 3 --warmup 3 --no-k1
step "raft / ssi N=2" bash -c 'for w in MCraft_t4l3 MCssi_4x3; do timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 1 --no-k1 --workload $w; done'
tail -3 "$LOG"
